#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, then the two ncu passes of B200_PROFILING.md on the same bench command.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
# Everything lands in gpurun_out/<tag>_*; scripts/summarize_profiles.py turns it into profiles/<tag>_*.md here.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -2 $out/${tag}_bench.err; cat $out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 2 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err; cat $out/${tag}_bench_ref.json
timeout 100 python scripts/profile_step.py router > $out/${tag}_prof_router.txt 2>&1
timeout 100 python scripts/profile_step.py competition > $out/${tag}_prof_comp.txt 2>&1
python bench.py --steps 2 --warmup 3 > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 > $out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 30 -c 6 -o $out/${tag}_gemm_full \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu2.log 2>&1
python bench.py --steps 2 --warmup 3 > $out/${tag}_plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'gather_rows|combine|scatter_reduce|act_bwd|router|route_|compete|affinity|diversity' -c 300 --csv \
    --log-file $out/${tag}_hbm.csv python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu3.log 2>&1
ls -la $out | tail -20
