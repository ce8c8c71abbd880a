"""Turn the raw ncu output of scripts/gpu_round.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [kernel-regex for the --set full page, default grouped_gemm]

Writes profiles/<tag>_launches.md (share of device time per kernel, from the launch list), profiles/<tag>_gemm_ncu_full.md
(one row per captured launch with the counters the roofline discussion uses) and profiles/gemm_traffic.json
(DRAM bytes per launch of the dominant kernel, read by bench.py for `roofline.traffic`).  Runs here (no GPU needed).
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__cycles_elapsed.avg.per_second",
]


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0][:110]


def launches(tag: str, cmd: str):
    path = OUT / f"{tag}_launches.csv"
    text = path.read_text()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(lambda: [0.0, 0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        agg[k][0] += us
        agg[k][1] += 1
    total = sum(v[0] for v in agg.values())
    n = sum(v[1] for v in agg.values())
    lines = [f"# {tag} -- ncu launch list of `{cmd}`", "",
             "Command (B200, after the same command exited 0 without ncu):", "",
             f"    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/{tag}_launches.csv {cmd}",
             "", f"{n} kernel launches, {total / 1e3:.1f} ms of summed device time (cold-cache, serialised: compare shares, not absolutes).",
             "The run covers the router-step, competition-step and end-to-end timed regions plus their warm-ups.", "",
             "| share | total us | launches | kernel |", "|---:|---:|---:|---|"]
    ours = 0.0
    gemm = 0.0
    for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        lines.append(f"| {100 * us / total:.1f}% | {us:.0f} | {c} | `{k}` |")
    for k, (us, c) in agg.items():
        if k.startswith("csmoe::"):
            ours += us
        if "grouped_gemm" in k:
            gemm += us
    lines += ["", f"libcsmoe kernels: **{100 * ours / total:.1f}%** of device time; grouped GEMM (tcgen05): **{100 * gemm / total:.1f}%**."]
    (PROF / f"{tag}_launches.md").write_text("\n".join(lines) + "\n")
    print(f"wrote profiles/{tag}_launches.md ({n} launches, GEMM share {100 * gemm / total:.1f}%)")


def full(tag: str, cmd: str, pattern: str):
    rep = OUT / f"{tag}_gemm_full.ncu-rep"
    if not rep.exists():
        print("no", rep)
        return
    r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    cols = ["Kernel Name"] + [m for m in METRICS if m in idx]
    lines = [f"# {tag} -- `ncu --set full` of the grouped GEMM launches of one router step", "",
             f"Command: `ncu --set full --clock-control none --import-source on -k regex:{pattern} -s 30 -c 6 -o gpurun_out/{tag}_gemm_full {cmd}`",
             "", "Template arguments: pair kernel `<MODE, B_MN>`, single-CTA kernel `<MODE, B_MN, BN>`; MODE 0 = ROWS (fwd / dgrad), 1 = REDUCE (wgrad).",
             "", "| " + " | ".join(f"{c} [{units[idx[c]]}]" if c in idx and units[idx[c]] else c for c in cols) + " |",
             "|" + "---|" * len(cols)]
    traffic = []
    for d in data:
        vals = []
        for c in cols:
            v = d[idx[c]]
            vals.append(short(v)[:60] if c == "Kernel Name" else v)
        lines.append("| " + " | ".join(vals) + " |")
        try:
            def tobytes(metric):
                v, u = float(d[idx[metric]].replace(",", "")), units[idx[metric]].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            traffic.append(tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))
        except Exception:
            pass
    (PROF / f"{tag}_gemm_ncu_full.md").write_text("\n".join(lines) + "\n")
    if traffic:
        (PROF / "gemm_traffic.json").write_text(json.dumps({
            "source": f"profiles/{tag}_gemm_ncu_full.md (dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full)",
            "per_launch_bytes": traffic, "dram_bytes_per_launch": sum(traffic) / len(traffic)}, indent=1) + "\n")
    print(f"wrote profiles/{tag}_gemm_ncu_full.md ({len(data)} launches)")


def hbm(tag: str, cmd: str):
    """Third ncu pass of gpu_round.sh: DRAM byte counters of the bandwidth-bound kernels (permute, combine, activation
    backward, router) inside the bench step -> profiles/<tag>_hbm_ncu.md (achieved GB/s = DRAM bytes / kernel time)."""
    path = OUT / f"{tag}_hbm.csv"
    if not path.exists():
        print("no", path)
        return
    text = path.read_text()
    rows = list(csv.DictReader(io.StringIO(text[text.find('"ID"'):])))
    per = defaultdict(dict)
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = (r.get("Metric Unit") or "").lower()
        scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "usecond": 1e-6, "us": 1e-6, "nsecond": 1e-9,
                 "ns": 1e-9, "msecond": 1e-3, "ms": 1e-3, "%": 1.0}.get(u, 1.0)
        per[(r["ID"], short(r["Kernel Name"]))][r["Metric Name"]] = v * scale
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for (_, k), m in per.items():
        a = agg[k]
        a[0] += m.get("gpu__time_duration.sum", 0.0)
        a[1] += m.get("dram__bytes_read.sum", 0.0)
        a[2] += m.get("dram__bytes_write.sum", 0.0)
        a[3] += 1
    lines = [f"# {tag} -- DRAM traffic of the bandwidth-bound kernels inside the bench step (ncu)", "",
             "Command (B200, after the same command exited 0 without ncu):", "",
             f"    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
             f"-k regex:'gather_rows|combine|scatter_reduce|act_bwd|router|route_|compete|affinity|diversity' --csv "
             f"--log-file gpurun_out/{tag}_hbm.csv {cmd}", "",
             "Per-launch averages; GB/s = (DRAM read + write) / kernel time, against the measured copy bandwidth 6555 GB/s "
             "(cold-cache, serialised launches: inside the step the producer's output is partly L2 resident, so DRAM bytes can "
             "be below the algorithmic bytes).", "",
             "| kernel | launches | us / launch | DRAM read MB | DRAM write MB | GB/s | % of 6555 |", "|---|---:|---:|---:|---:|---:|---:|"]
    for k, (t, rd, wr, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if t <= 0:
            continue
        gbs = (rd + wr) / t / 1e9
        lines.append(f"| `{k[:70]}` | {n} | {1e6 * t / n:.1f} | {rd / n / 1e6:.1f} | {wr / n / 1e6:.1f} | {gbs:.0f} | {100 * gbs / 6554.9:.1f} |")
    (PROF / f"{tag}_hbm_ncu.md").write_text("\n".join(lines) + "\n")
    print(f"wrote profiles/{tag}_hbm_ncu.md ({len(per)} launches)")


if __name__ == "__main__":
    tag = sys.argv[1]
    pattern = sys.argv[2] if len(sys.argv) > 2 else "grouped_gemm"
    cmd = "python bench.py --steps 2 --warmup 3"
    PROF.mkdir(exist_ok=True)
    launches(tag, cmd)
    full(tag, cmd, pattern)
    hbm(tag, cmd)
