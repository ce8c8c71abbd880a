#!/bin/bash
# Round-2 profiling call (one gpurun call; every ncu pass follows a plain run of the same command that exited 0):
#   1. launch list of the bench command (headline regions only: --sections 0)
#   2. --set full of the six grouped-GEMM launches of the bench step
#   3. --set full of the fused sigma-MoE kernels at the C4 shape (scripts/sigma_prof.py: fwd, bwd, wgrad x2)
#   4. launch list of one C4 router + competition step (scripts/config_sweep.py --only C4)
# Usage: bash scripts/gpu_round2.sh <tag>;  scripts/summarize_profiles.py <tag> writes profiles/<tag>_*.md here.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --sections 0"
$B > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu1.log 2>&1
$B > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 30 -c 6 -o $out/${tag}_gemm_full $B > $out/${tag}_ncu2.log 2>&1
python scripts/sigma_prof.py 2 > $out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sigma -s 4 -c 4 -o $out/${tag}_sigma_full python scripts/sigma_prof.py 2 > $out/${tag}_ncu3.log 2>&1
S="python scripts/config_sweep.py --only C4 --steps 1 --warmup 2"
$S > $out/${tag}_plain4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/${tag}_c4_launches.csv $S > $out/${tag}_ncu4.log 2>&1
ls -la $out | tail -12
