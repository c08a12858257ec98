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


def _build_one(name):
    """Compile one reference module in this process (torch's JIT recipe, in-tree build directory)."""
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "2")
    from torch.utils.cpp_extension import load
    spec = SPECS[name]
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    src = os.path.join(REF, spec["dir"], "src")
    load(name=name, extra_cflags=CXX, extra_cuda_cflags=NVCC + spec.get("nvcc", []),
         extra_include_paths=[os.path.join(REF, spec["dir"], p) for p in spec.get("inc", [])],
         sources=[os.path.join(src, f) for f in spec["srcs"]],
         build_directory=bdir, is_python_module=False, verbose=False)


def build(names=None):
    """Build the missing modules, one child process each, all at once: every module is a single
    long nvcc translation unit (2.5-5.5 min), so the four in parallel cost what `_ffmlp` costs alone."""
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: skipping (prebuilt oracle/_ref is used if it exists)")
        return False
    import subprocess
    todo = []
    for name in (names or list(SPECS)):
        if os.path.exists(os.path.join(OUT, name, name + ".so")):
            print(f"[build_ref] {name}: already built")
        else:
            todo.append(name)
    procs = [(n, subprocess.Popen([sys.executable, os.path.abspath(__file__), "--one", n],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for n in todo]
    ok = True
    for n, p in procs:
        out, _ = p.communicate()
        if p.returncode == 0:
            print(f"[build_ref] {n}: built -> {OUT}/{n}/{n}.so")
        else:
            ok = False
            print(f"[build_ref] {n}: FAILED:\n{out[-2000:]}")
    return ok


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        _build_one(sys.argv[2])
        sys.exit(0)
    sys.exit(0 if build(sys.argv[1:] or None) else 1)
