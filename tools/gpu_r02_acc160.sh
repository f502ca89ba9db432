#!/bin/bash
# round 2 experiment: may a 160-column tcgen05.mma accumulator start at TMEM column 160 (32-aligned, not 256-aligned)?
# (mobi_b200/_build/libmobi_exp.so = the library with gemm2.cu compiled with -DG2_ACC160)
mkdir -p gpurun_out
MOBI_B200_LIB=$PWD/mobi_b200/_build/libmobi_exp.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm_plain or conv_implicit or quad or colstats" 2>&1 | tail -8
MOBI_B200_LIB=$PWD/mobi_b200/_build/libmobi_exp.so KB_GEMM_TILE=160 timeout 100 python tools/kbench.py conv 2>&1 | grep "^conv"
