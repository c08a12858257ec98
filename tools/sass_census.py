#!/usr/bin/env python
"""Which Blackwell instructions the built library contains, per kernel family: `cuobjdump -sass` of enerf_b200/libenerf_b200.so, the
mnemonics that prove tcgen05 / TMEM / TMA use (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads /
stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations) next to the legacy ones (HMMA = mma.sync) and the reductions (REDG).
Runs without a GPU.    python tools/sass_census.py > profiles/<tag>_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "enerf_b200", "libenerf_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FENCE.VIEW.ASYNC", "REDG", "ATOMG", "HMMA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    n_inst = collections.Counter()
    archs, cur = set(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            archs.add(m.group(1))
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            n_inst[cur] += 1
            for op in OPS:
                if re.search(r"\b" + re.escape(op), line):
                    counts[cur][op] += 1
    names = list(n_inst)
    demangled = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    agg = collections.defaultdict(collections.Counter)
    inst, kernels = collections.Counter(), collections.Counter()
    for mangled, d in zip(names, demangled):
        base = re.sub(r"<.*", "", d.split("(")[0]).replace("void ", "")
        agg[base].update(counts[mangled])
        inst[base] += n_inst[mangled]
        kernels[base] += 1
    total = collections.Counter()
    for c in agg.values():
        total.update(c)
    used = [op for op in OPS if total[op]]
    print(f"# SASS census of `{os.path.relpath(lib)}` ({os.path.getsize(lib) / 2**20:.1f} MiB; code objects: {', '.join(sorted(archs))}; "
          f"{sum(kernels.values())} kernels in {len(kernels)} families, {sum(inst.values())} instructions)\n")
    print("`tools/sass_census.py`.  Totals: " + ", ".join(f"{total[op]} `{op}`" for op in used) + "\n")
    print("| kernel family (instantiations) | SASS instructions | " + " | ".join(f"`{op}`" for op in used) + " |")
    print("|---|---|" + "---|" * len(used))
    for base in sorted(inst, key=lambda b: -inst[b]):
        print(f"| `{base}` ({kernels[base]}) | {inst[base]} | " + " | ".join(str(agg[base][op]) if agg[base][op] else "" for op in used) + " |")


if __name__ == "__main__":
    main()
