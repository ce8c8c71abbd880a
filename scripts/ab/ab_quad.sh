#!/bin/bash
out=gpurun_out/ab_quad.txt
: > $out
CSMOE_VARIANT_CHILD=1 CSMOE_GEMM_QUAD=3 CSMOE_GEMM_WIDE=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm and not variants" 2>&1 | tail -15 >> $out
for cfg in "CSMOE_GEMM_QUAD=0" "CSMOE_GEMM_QUAD=3" "CSMOE_GEMM_QUAD=3 CSMOE_GEMM_WIDE=0"; do
  echo "== $cfg" >> $out
  env $cfg timeout 200 python scripts/gemm_bench.py 30 2>&1 | grep -v cuBLAS >> $out
done
echo "== stats QUAD=3 WIDE=0" >> $out
CSMOE_GEMM_QUAD=3 CSMOE_GEMM_WIDE=0 CSMOE_GEMM_STATS=1 timeout 200 python scripts/gemm_bench.py 4 2>&1 | grep "stats" | awk 'NR%7==0' >> $out
cat $out
