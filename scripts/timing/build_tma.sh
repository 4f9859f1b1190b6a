#!/bin/bash
# builds scripts/timing/libeat_tma_timing.so: pw_tma.cu with the per-role cycle accounting compiled in
cd "$(dirname "$0")/../.."
python -m efficientat_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DEAT_TMA_TIMING \
  -c efficientat_b200/csrc/pw_tma.cu -o /tmp/pw_tma_timing.o
nvcc -shared -o scripts/timing/libeat_tma_timing.so /tmp/pw_tma_timing.o efficientat_b200/build/api.o -gencode arch=compute_100a,code=sm_100a
