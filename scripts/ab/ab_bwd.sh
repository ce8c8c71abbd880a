#!/bin/bash
out=gpurun_out/ab_bwd.txt
: > $out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -5 >> $out
for cfg in "CSMOE_GEMM_BWD_TMA=0" "CSMOE_GEMM_BWD_TMA=1"; do
  echo "== $cfg" >> $out
  env $cfg timeout 200 python scripts/gemm_bench.py 30 2>&1 | grep -E "dgrad2|act_bwd" >> $out
done
CSMOE_GEMM_STATS=1 timeout 200 python scripts/gemm_bench.py 4 2>&1 | grep "stats" | grep "epi=3" | tail -1 >> $out
for f in 0 1; do
  echo "== bench CSMOE_FUSE_EPILOGUE_BWD=$f" >> $out
  CSMOE_FUSE_EPILOGUE_BWD=$f timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['eager']['ms_per_step'], d['competition']['ms_per_step'], d['roofline']['frac'])" >> $out
done
cat $out
