#!/bin/bash
out=gpurun_out/ab_tma2.txt
: > $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -3 >> $out
for cfg in "CSMOE_GEMM_RASTER=1" "CSMOE_GEMM_RASTER=1 CSMOE_GEMM_WIDE=15"; do
  echo "== $cfg" >> $out
  env $cfg timeout 200 python scripts/gemm_bench.py 30 2>&1 | grep -v cuBLAS >> $out
done
echo "== stats: RASTER=1 WIDE=15" >> $out
CSMOE_GEMM_RASTER=1 CSMOE_GEMM_WIDE=15 CSMOE_GEMM_STATS=1 timeout 200 python scripts/gemm_bench.py 4 2>&1 | grep "stats" | awk 'NR%7==0' >> $out
echo "== stats: RASTER=1 pair" >> $out
CSMOE_GEMM_RASTER=1 CSMOE_GEMM_WIDE=0 CSMOE_GEMM_STATS=1 timeout 200 python scripts/gemm_bench.py 4 2>&1 | grep "stats" | awk 'NR%7==0' >> $out
echo "== siglip RASTER=1" >> $out
CSMOE_GEMM_RASTER=1 timeout 200 python scripts/gemm_bench.py 30 siglip 2>&1 | grep -v cuBLAS >> $out
cat $out
