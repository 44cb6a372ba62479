#!/bin/bash
# First-call times of the reference's main.cpp on the GPU (bin/spmm_main, every function called once) under the CUDA module
# loading modes, with the per-phase timings of the host path (SPMM_HOST_TIMING).
python tools/write_cfg_mtx.py cfg2 /tmp/cfg2.mtx > /dev/null 2>&1
M=./sparsematrixmultiplicationmpi_b200/bin/spmm_main
for mode in ${MODES:-unset LAZY EAGER LAZY LAZY}; do for np in ${NPS:-1 2}; do
  echo "== mode=$mode np=$np"
  s=$(date +%s%N)
  if [ $mode = unset ]; then SPMM_HOST_TIMING=1 $M -np $np 64 /tmp/cfg2.mtx 2>&1 | grep -E "Execution time|row-wise|spmm entry\] shard" | grep -v PETSc
  else SPMM_HOST_TIMING=1 CUDA_MODULE_LOADING=$mode $M -np $np 64 /tmp/cfg2.mtx 2>&1 | grep -E "Execution time|row-wise|spmm entry\] shard" | grep -v PETSc; fi
  e=$(date +%s%N); echo " wall $(( (e - s) / 1000000 )) ms"
done; done
