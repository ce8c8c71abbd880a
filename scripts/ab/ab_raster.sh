#!/bin/bash
out=gpurun_out/ab_raster.txt
: > $out
run() { echo "== $*" >> $out; env "$@" CSMOE_GEMM_WIDE=15 timeout 200 python scripts/gemm_bench.py 20 2>&1 | grep -E "fwd1 plain|fwd2 |dgrad1|dgrad2  " >> $out; }
run CSMOE_GEMM_RASTER=0
run CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=8
run CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=4
run CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=16
run CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=37
run CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=74
for r in "CSMOE_GEMM_RASTER=0" "CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=8" "CSMOE_GEMM_RASTER=1 CSMOE_GEMM_BAND=37"; do
  echo "== ncu $r" >> $out
  env $r CSMOE_GEMM_WIDE=15 ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:grouped_gemm -s 2 -c 1 python scripts/gemm_one.py 2>&1 | grep -E "gpu__time|lts__t|dram__|tensor" >> $out
done
cat $out
