#!/bin/bash
# Instrumented developer builds of the tcgen05 attention forward (csrc/flash_tc.cu under -DFT_PROFILE: per-phase cycle counters,
# read by tools/exp/ft_profile.py).  libevk_ftprof.so: mbarrier.try_wait; libevk_ftprof_tw.so: mbarrier.test_wait polling.
set -e
cd "$(dirname "$0")/../.."
PKG=easevoice-trainer_b200
python -c "import sys; sys.path.insert(0, '$PKG'); import build; build.build()"
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DFT_PROFILE"
nvcc $FL -c $PKG/csrc/flash_tc.cu -o tools/exp/flash_tc_prof.o
nvcc $FL -DFT_TESTWAIT -c $PKG/csrc/flash_tc.cu -o tools/exp/flash_tc_prof_tw.o
OBJS=$(ls $PKG/build/*.o | grep -v flash_tc.o)
nvcc -shared -o tools/exp/libevk_ftprof.so $OBJS tools/exp/flash_tc_prof.o -gencode arch=compute_100a,code=sm_100a -lcudart
nvcc -shared -o tools/exp/libevk_ftprof_tw.so $OBJS tools/exp/flash_tc_prof_tw.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built tools/exp/libevk_ftprof.so tools/exp/libevk_ftprof_tw.so
