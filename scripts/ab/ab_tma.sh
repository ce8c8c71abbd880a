#!/bin/bash
# A/B of the TMA-store epilogue on the bench-shape GEMMs (gemm_bench.py), run under gpurun.
out=gpurun_out/ab_tma.txt
: > $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5 >> $out
for cfg in "CSMOE_GEMM_TMA=0" "CSMOE_GEMM_TMA=3" "CSMOE_GEMM_TMA=3 CSMOE_GEMM_WIDE=15" "CSMOE_GEMM_TMA=0 CSMOE_GEMM_WIDE=15"; do
  echo "== $cfg" >> $out
  env $cfg timeout 200 python scripts/gemm_bench.py 30 2>&1 | grep -v cuBLAS >> $out
done
echo "== stats: TMA=3 WIDE=15" >> $out
CSMOE_GEMM_TMA=3 CSMOE_GEMM_WIDE=15 CSMOE_GEMM_STATS=1 timeout 200 python scripts/gemm_bench.py 4 2>&1 | grep "stats" | awk 'NR%7==0' >> $out
echo "== siglip TMA=0" >> $out
CSMOE_GEMM_TMA=0 timeout 200 python scripts/gemm_bench.py 30 siglip 2>&1 | grep -v cuBLAS >> $out
echo "== siglip TMA=3" >> $out
CSMOE_GEMM_TMA=3 timeout 200 python scripts/gemm_bench.py 30 siglip 2>&1 | grep -v cuBLAS >> $out
cat $out
