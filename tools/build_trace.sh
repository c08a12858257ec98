#!/bin/bash
# Build tools/bin/libenerf_b200_trace.so: the same library with -DENERF_TC_TRACE (cycle stamps in the recompute-backward kernel).
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin/trace_obj
for f in enerf_b200/csrc/*.cu; do
  o=tools/bin/trace_obj/$(basename ${f%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DENERF_TC_TRACE -c $f -o $o &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/bin/libenerf_b200_trace.so tools/bin/trace_obj/*.o
echo built tools/bin/libenerf_b200_trace.so
