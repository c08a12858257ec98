"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The four pybind11 modules built here
(`_raymarching`, `_gridencoder`, `_shencoder`, `_ffmlp`) are the reference's own
kernels, compiled from the sources where they lie under /root/reference (nothing
is copied into this repository).  They are the GPU-side oracle: the `-m gpu`
parity tests load them from oracle/_ref/ (when present) and compare this repo's
kernels against them on identical seeded inputs.  Nothing in the product path
(`enerf_b200/`) may import them.

Recipe = the reference's own JIT recipe (`*/backend.py`) with one forced change:
`-std=c++14` -> `-std=c++17` (torch 2.11 headers `#error` on C++14) and the arch
pinned to sm_100a.  The built .so files are git-ignored but travel to the GPU box.

Usage:  python oracle/build_ref.py [name ...]      (default: all four)
"""
import os
import sys

REF = os.environ.get("ENERF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

NVCC = ['-O3', '-std=c++17',
        '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
        '-U__CUDA_NO_HALF2_OPERATORS__']
CXX = ['-O3', '-std=c++17']

SPECS = {
    "_raymarching": dict(dir="raymarching", srcs=["raymarching.cu", "bindings.cpp"]),
    "_gridencoder": dict(dir="gridencoder", srcs=["gridencoder.cu", "bindings.cpp"]),
    "_shencoder": dict(dir="shencoder", srcs=["shencoder.cu", "bindings.cpp"]),
    "_ffmlp": dict(dir="ffmlp", srcs=["ffmlp.cu", "bindings.cpp"],
                   nvcc=['--expt-extended-lambda', '--expt-relaxed-constexpr',
                         '-Xcompiler=-mf16c', '-Xcompiler=-Wno-float-conversion',
                         '-Xcompiler=-fno-strict-aliasing'],
                   inc=["dependencies/cutlass/include", "dependencies/cutlass/tools/util/include"]),
}


def build(names=None):
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: skipping (prebuilt oracle/_ref is used if it exists)")
        return False
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    ok = True
    for name in (names or list(SPECS)):
        spec = SPECS[name]
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        if os.path.exists(os.path.join(bdir, name + ".so")):
            print(f"[build_ref] {name}: already built")
            continue
        src = os.path.join(REF, spec["dir"], "src")
        try:
            load(name=name, extra_cflags=CXX, extra_cuda_cflags=NVCC + spec.get("nvcc", []),
                 extra_include_paths=[os.path.join(REF, spec["dir"], p) for p in spec.get("inc", [])],
                 sources=[os.path.join(src, f) for f in spec["srcs"]],
                 build_directory=bdir, is_python_module=False, verbose=False)
            print(f"[build_ref] {name}: built -> {bdir}/{name}.so")
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"[build_ref] {name}: FAILED: {e}")
    return ok


if __name__ == "__main__":
    sys.exit(0 if build(sys.argv[1:] or None) else 1)
