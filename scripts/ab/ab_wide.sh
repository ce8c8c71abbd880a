#!/bin/bash
out=gpurun_out/ab_wide.txt
: > $out
run() { echo "== $*" >> $out; env "$@" CSMOE_GEMM_WIDE=15 timeout 200 python scripts/gemm_bench.py 20 2>&1 | grep -E "fwd1 plain|fwd2 |dgrad1|wgrad1" >> $out; }
run CSMOE_GEMM_DBG=0
run CSMOE_GEMM_DBG=1
run CSMOE_GEMM_DBG=2
run CSMOE_GEMM_BAND=2
run CSMOE_GEMM_BAND=4
run CSMOE_GEMM_BAND=16
run CSMOE_GEMM_BAND=64
cat $out
