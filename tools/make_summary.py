"""Rebuild the head (bench lines, kernel table, launch list, ncu table) of profiles/r01_summary.md from
profiles/r01_bench_*.json and the ncu outputs in gpurun_out/; the hand-written "## Readings" section is kept.

    python tools/make_summary.py
"""
import json
import subprocess
import sys

tabs = subprocess.run([sys.executable, "tools/summarize_profiles.py", "gpurun_out/launches_r01b.csv",
                       "gpurun_out/prof_r01b_k_tridiag_packed.ncu-rep", "gpurun_out/prof_r01b_k_tql_smem.ncu-rep",
                       "gpurun_out/prof_r01b_k_replay.ncu-rep", "gpurun_out/prof_r01b_k_ne_dmma3.ncu-rep"],
                      capture_output=True, text=True, check=True).stdout.split("\n\n")
d = json.load(open("profiles/r01_bench_default.json"))
d2 = json.load(open("profiles/r01_bench_2gpu.json"))
d8 = json.load(open("profiles/r01_bench_8gpu.json"))
ref = json.load(open("profiles/r01_bench_reference.json"))
launch = [l for l in tabs[0].split("\n") if not l.startswith("| at::")]
k = d["kernels"]
rows = []
for name in ["normal_eq", "tridiag", "tql", "apply", "chi2", "covariance", "misc"]:
    v = k[name]
    tf, fr = v.get("tflops"), v.get("frac_of_fp64_peak")
    rows.append(f"| {name} | {v['ms_per_step']:.1f} | {v['launches_per_step']:.0f} | {100 * v['share']:.1f} % | "
                f"{('%.2f' % tf) if tf else '-'} | {('%.1f %%' % (100 * fr)) if fr else '-'} |")
est = d["estimate"]
old = open("profiles/r01_summary.md").read()
readings = old[old.index("## Readings"):]
history = old[old.index("Progress over the round"):old.index("## Launch list")]
head = f"""# Round 1 profile summary (B200 sm_100a, CUDA 12.9, driver 580)

All captures: the same command was first run without ncu and exited 0 (`tools/profile_round.sh`; tables made by
`tools/summarize_profiles.py`, this file by `tools/make_summary.py`).  State of the tree: end of round 1 (packed
tridiagonalisation, sparse-lane QL / replay, normal equations v3, staged eigenvector kernel).

## Bench lines (`python bench.py` -> r01_bench_default.json; `torchrun ... bench.py --gpus 2` -> r01_bench_2gpu.json;
`python bench.py --impl reference` -> r01_bench_reference.json)

* fitted records/s, device-resident inputs: **{d['value']:.0f}** ({d['ms_per_step']:.0f} ms per 10 000-record step, covariance included); 2 GPUs: **{d2['value']:.0f}** (weak scaling, {d2['ms_per_step']:.0f} ms per step, NCCL all-gather of the coefficients inside the timed region)
* 8 GPUs (`r01_bench_8gpu.json`, 2 steps): **{d8['value']:.0f}** records/s device-resident ({d8['ms_per_step']:.0f} ms per step: {100 * d8['value'] / (8 * d['value']):.0f} % of 8 x the single-GPU rate), {d8['e2e']['value']:.0f} end to end (eight ranks share 16 host cores and the host's PCIe / memory bandwidth for 13 GB of covariance per step)
* end to end (pinned host buffers -> numpy): **{d['e2e']['value']:.0f}** records/s, H2D {d['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB + D2H {d['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB per step; 2 GPUs: {d2['e2e']['value']:.0f}
* reference algorithm on the box's host cores ({d['cpu_baseline']['cores']} processes): {d['cpu_baseline']['value']:.3f} records/s in the same run, {ref['value']:.3f} in the reference arm
* clocks during the timed region: {d['clocks']}
* {d['fit']['eigen_systems_per_step']} eigen-systems per step ({d['fit']['eigen_systems_per_step'] / 10000:.1f} per record); status histogram {d['fit']['status_histogram']}
* Estimate: {est['single_record']['value']:.3e} points/s (one record, {est['single_record']['points']} points, hull mask on); {est['records64']['value']:.3e} (point,record)/s for 64 records; {est['records512']['value']:.3e} for 512 records on the DMMA path ({100 * est['records512']['frac_of_fp64_peak']:.0f} % of FP64 peak); CPU port {est['cpu_baseline']['value']:.0f} points/s on one core, masks identical

| kernel kind (CUDA events in the timed region) | ms per step | launches per step | share | TFLOP/s | of FP64 peak |
|---|---|---|---|---|---|
""" + "\n".join(rows) + f"""

FP64 peaks measured in the same run: {d['roofline']['fp64_peaks_tflops']} TFLOP/s.

""" + history + """## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, 300 records, file r01_launches_records300.csv)

""" + "\n".join(launch) + """

At 300 records every Brent round is a few hundred systems, so the thread-per-system kernels (k_tql_smem / k_tql_single,
k_replay) run one partially filled wave per launch and weigh far more than in the 10 k-record step above, where
k_tridiag_packed is 62 %.

## `ncu --set full`, one launch each of the table phase at 2000 records (28 000 systems per launch)

""" + tabs[1].rstrip() + "\n\n"
open("profiles/r01_summary.md", "w").write(head + readings)
print("profiles/r01_summary.md rewritten")
