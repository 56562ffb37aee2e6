python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; tail -4 gpurun_out/r02f_pytest.log
python tools/time_solver.py 28416 144 2>&1 | tail -1
python tools/time_solver.py 600 144 2>&1 | tail -1
python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02f_plain.log 2> gpurun_out/r02f_plain.err || { tail -5 gpurun_out/r02f_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02f_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/launches_r02f.csv")))
h=next(i for i,r in enumerate(rows) if "Kernel Name" in r)
H=rows[h]; ik=H.index("Kernel Name"); iv=H.index("Metric Value")
t=collections.defaultdict(lambda:[0,0.0])
for r in rows[h+1:]:
    if len(r)>iv:
        n=r[ik].split("(")[0][:40]; t[n][0]+=1; t[n][1]+=float(r[iv].replace(",",""))/1e6
tot=sum(v[1] for v in t.values())
for k,v in sorted(t.items(), key=lambda x:-x[1][1])[:22]: print(f"{k:42s} {v[0]:5d} {v[1]:9.2f} ms {100*v[1]/tot:5.1f}%")
PY
