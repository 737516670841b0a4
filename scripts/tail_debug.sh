#!/bin/bash
# build container: rebuild finalize.cu with in-kernel phase stamps, run on the GPU box, restore the normal build
set -e
cd /root/repo/pyvbmc_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DVBMC_TAIL_DEBUG -c finalize.cu -o finalize.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libvbmc_b200.so capi.o entmc.o entmc_tc.o gplj.o gpvar.o entlb.o finalize.o peaks.o -lcudart_static -lpthread -ldl -lrt
cd /root/repo
gpurun --timeout 300 -- 'python scripts/run_negelcbo.py C3 4 2>&1 | tail -9' 2>&1 | tail -12
cd /root/repo/pyvbmc_b200/csrc && rm finalize.o && make -j8 > /dev/null 2>&1
