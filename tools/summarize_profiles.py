"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into markdown tables for profiles/.

    python tools/summarize_profiles.py gpurun_out/launches_r01b.csv gpurun_out/prof_r01b_*.ncu-rep
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def launch_table(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr) or r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
        n, t = agg.get(name, (0, 0.0))
        unit = r[hdr.index("Metric Unit")]
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {name} | {n} | {t:.2f} | {100 * t / tot:.1f} % |")
    return "\n".join(out)


WANT = [("gpu__time_duration.sum", "duration ms", 1e-6), ("launch__registers_per_thread", "regs", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe %", 1),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts", 1),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts", 1),
        ("dram__bytes_read.sum", "DRAM read MB", 1e-6), ("dram__bytes_write.sum", "DRAM write MB", 1e-6),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes / instr", 1)]


def full_table(reps):
    out = ["| kernel (grid) | " + " | ".join(w[1] for w in WANT) + " |", "|---|" + "---|" * len(WANT)]
    for rep in reps:
        if rep.endswith(".csv"):      # raw page exported on the GPU box (tools/profile_round2_ncu.sh)
            txt = open(rep).read()
        else:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                                 text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr = rows[0]
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("<unnamed>::", "").replace("void ", "")
            cells = []
            for key, _, scale in WANT:
                if key in hdr:
                    try:
                        v = float(r[hdr.index(key)].replace(",", "")) * scale
                        cells.append(f"{v:.4g}")
                    except ValueError:
                        cells.append(r[hdr.index(key)])
                else:
                    cells.append("-")
            out.append(f"| {name} {r[hdr.index('Grid Size')]} | " + " | ".join(cells) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    print(launch_table(sys.argv[1]))
    print()
    print(full_table(sys.argv[2:]))
