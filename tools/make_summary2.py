"""profiles/r02_summary.md + profiles/r02_traffic.json from the evidence run of tools/profile_round2.sh
(profiles/r02_bench_*.json, gpurun_out/r02_launches_records2000.csv, gpurun_out/prof_r02_*.ncu-rep).
The hand-written "## Readings" section of an existing file is kept.

    python tools/make_summary2.py
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, "tools")
from summarize_profiles import full_table, launch_table

P = "profiles"


def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


d = load("r02_bench_default.json")
ref = load("r02_bench_reference.json")
out = ["# Round 2 profile summary (B200 sm_100a, CUDA 12.9, driver 580)", "",
       "Every command was first run without a profiler and exited 0 (`tools/profile_round2.sh`); the bench lines below are",
       "those runs.  Tables by `tools/summarize_profiles.py`, this file by `tools/make_summary2.py`.", "",
       "## Bench lines (`profiles/r02_bench_*.json`)", ""]


def one(tag, d, what):
    if d is None:
        return
    e = d.get("e2e") or {}
    out.append(f"* **{tag}** ({what}): {d['value']:.4g} {d['unit']}, {d.get('ms_per_step', 0):.4g} ms per step, "
               f"end to end {e.get('value', float('nan')):.4g}" + (f", clocks {d['clocks']}" if d.get("clocks") else ""))


one("default", d, "`python bench.py`: C2, 10 000 records, one GPU")
one("reference arm", ref, "`python bench.py --impl reference`: the reference algorithm, 16 host processes")
one("C1", load("r02_bench_c1.json"), "`--config c1`: 51 x 14 gates, 300 records")
one("C3", load("r02_bench_c3.json"), "`--config c3`: N = 500, 51 x 200 gates")
one("C4", load("r02_bench_c4.json"), "`--config c4`: Estimate, 512^3 grid x 1000 records, 96 of 1024 tiles spread over the grid")
one("C5", load("r02_bench_c5.json"), "`--config c5 --records 64`: leave-beam-out, 64 x 51 refits")
one("NaN-heavy mix", load("r02_bench_mix.json"), "`--noise-scale 1.0 --signal-terms 5`: the round-1 workload (15 % NaN records)")
one("2 GPUs weak", load("r02_bench_2gpu_weak.json"), "torchrun, 10 000 records per GPU")
one("2 GPUs strong", load("r02_bench_2gpu_strong.json"), "torchrun `--scaling strong --records 10000`")
one("2 GPUs C4", load("r02_bench_2gpu_c4.json"), "torchrun `--config c4 --tiles 32`, point-sharded")
one("4 GPUs weak", load("r02_bench_4gpu_weak.json"), "torchrun, 10 000 records per GPU, `--steps 2`")
one("4 GPUs C4", load("r02_bench_4gpu_c4.json"), "torchrun `--config c4 --tiles 32`")
one("4 GPUs strong", load("r02_bench_4gpu_strong.json"), "torchrun `--scaling strong --records 10000 --steps 2`")
one("8 GPUs weak", load("r02_bench_8gpu_weak.json"), "torchrun, 10 000 records per GPU, `--steps 2`")
one("8 GPUs strong", load("r02_bench_8gpu_strong.json"), "torchrun `--scaling strong --records 10000 --steps 2`")
one("8 GPUs C4", load("r02_bench_8gpu_c4.json"), "torchrun `--config c4 --tiles 32`")
if d:
    out += ["", f"* fit of the default run: {d.get('fit')}",
            f"* CPU baseline inside the default run: {d.get('cpu_baseline')}",
            f"* Estimate block of the default run: single record {d['estimate']['single_record']['value']:.4g} points/s "
            f"({d['estimate']['single_record']['inside_hull_fraction']:.2f} inside the hull)" if d.get("estimate") else ""]
    k = d["kernels"]
    out += ["", "| kernel kind (CUDA events in the timed region) | ms per step | launches per step | share | algorithmic rate | of its peak (FP64 37.1 TFLOP/s or HBM 6553 GB/s) |",
            "|---|---|---|---|---|---|"]
    for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        a, u, fr = v.get("achieved"), v.get("unit"), v.get("frac_of_peak")
        out.append(f"| {name} | {v['ms_per_step']:.1f} | {v['launches_per_step']:.0f} | {100 * (v['share'] or 0):.1f} % | "
                   f"{('%.4g %s' % (a, u)) if a else '-'} | {('%.1f %%' % (100 * fr)) if fr else '-'} |")
    r = d["roofline"]
    out += ["", f"Roofline block of the line: `{json.dumps({k2: r[k2] for k2 in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic') if k2 in r})}`",
            f"FP64 peaks measured in the same run: {r.get('fp64_peaks_tflops')} TFLOP/s; HBM {r.get('hbm_peak_gbs')} GB/s ({r.get('hbm_peak_source')})."]
    p = d.get("parity")
    if p:
        out += ["", "Parity block of the line (first 16 records of the benchmarked workload, GPU vs the reference algorithm): "
                f"`{json.dumps(p['gpu_vs_reference'])[:900]}`"]

lst0 = os.path.join(P, "r02_launches_records10000.csv")
if os.path.exists(lst0):
    out += ["", "## Launch list of the default workload (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
            "`bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-estimate`: 10 000 records, three passes; "
            "`profiles/r02_launches_records10000.csv`)", "",
            "The kernels' shares agree with the CUDA-event shares of the bench line above (stage 1 28.4 vs 28.5 %, QL 20.0 vs 20.3 %, "
            "chase 19.3 vs 19.4 %, replay 16.6 vs 16.5 %, chi2 5.5 vs 5.4 %, covariance 6.6 vs 6.6 %, normal equations 2.7 vs 2.7 %).", "",
            "\n".join(l for l in launch_table(lst0).split("\n")[:22] if not l.startswith("| at::") and "elementwise" not in l)]
lst = "gpurun_out/r02_launches_records2000.csv"
if os.path.exists(lst):
    out += ["", "## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
            "`bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline`; copy in "
            "`profiles/r02_launches_records2000.csv`)", "",
            "\n".join(l for l in launch_table(lst).split("\n") if not l.startswith("| at::") and "elementwise" not in l)]
reps = [f"gpurun_out/prof_r02_{k}.raw.csv" for k in ("k_band", "k_band_tail", "k_chase", "k_tql_smem", "k_replay_wave")]
reps = [r for r in reps if os.path.exists(r)]
if reps:
    out += ["", "## `ncu --set full` of the solver kernels (`tools/time_solver.py 8192 144`, one launch each)", "", full_table(reps)]
    traffic = {}
    for rep in reps:
        rows = list(csv.reader(io.StringIO(open(rep).read())))
        h = rows[0]
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("<unnamed>::", "").replace("void ", "")
            rd = float(r[h.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
            traffic[name] = {"systems_per_launch": 8192, "dram_bytes_read": rd, "dram_bytes_written": wr,
                             "dram_bytes_per_system": (rd + wr) / 8192}
    kind_of = {"k_band": "tridiag", "k_band_tail": "tridiag", "k_chase": "chase", "k_tql_smem": "tql", "k_replay_wave": "apply"}
    by_kind = {}
    for name, v in traffic.items():
        base = re.sub(r"<.*", "", name)
        if base in kind_of:
            by_kind[kind_of[base]] = by_kind.get(kind_of[base], 0.0) + v["dram_bytes_per_system"]
    json.dump({"capture": "profiles/r02_ncu_<kernel>.raw.csv (tools/time_solver.py 8192 144, ncu --set full, one launch each)",
               "dram_bytes_per_system": by_kind, "per_kernel": traffic}, open(os.path.join(P, "r02_traffic.json"), "w"), indent=1)
    out += ["", "DRAM traffic per eigen-system (`profiles/r02_traffic.json`): " +
            ", ".join(f"{k}: {v['dram_bytes_per_system'] / 1e3:.0f} KB" for k, v in traffic.items())]
path = os.path.join(P, "r02_summary.md")
readings = ""
if os.path.exists(path):
    old = open(path).read()
    if "## Readings" in old:
        readings = old[old.index("## Readings"):]
open(path, "w").write("\n".join(x for x in out if x is not None) + "\n\n" + readings)
print("wrote", path)
