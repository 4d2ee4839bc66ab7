#!/bin/bash
# Instrumented developer build: libevk_prof.so = the product objects with gemm_tma.cu recompiled under -DGT_PROFILE
# (per-CTA cycle counters of the producer / MMA / epilogue roles, read by tools/exp/gt_profile.py).
set -e
cd "$(dirname "$0")/../.."
PKG=easevoice-trainer_b200
python -c "import sys; sys.path.insert(0, '$PKG'); import build; build.build()"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DGT_PROFILE \
     -c $PKG/csrc/gemm_tma.cu -o tools/exp/gemm_tma_prof.o
OBJS=$(ls $PKG/build/*.o | grep -v gemm_tma.o)
nvcc -shared -o tools/exp/libevk_prof.so $OBJS tools/exp/gemm_tma_prof.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built tools/exp/libevk_prof.so
