#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
tail -6 gpurun_out/s5_pytest.log
python tools/tune_cutout.py > gpurun_out/s5_cutout_sweep.txt 2>&1; cat gpurun_out/s5_cutout_sweep.txt
timeout 600 python bench.py --no-cpu-baseline --no-parity-spot > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s5_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s5_bench.json"))
print("value %.0f e2e %.0f"%(d["value"],d["e2e"]["value"]), d["stage_ms_per_step"], d["clocks"])
PY
timeout 300 python bench.py --workload train --steps 4 > gpurun_out/s5_train_plain.json 2> gpurun_out/s5_train_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 4000 --csv --log-file gpurun_out/s5_train_launches.csv python bench.py --workload train --steps 4 > gpurun_out/s5_train_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open("gpurun_out/s5_train_launches.csv") if l.startswith('"')))
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); iu=hdr.index("Metric Unit")
agg=collections.defaultdict(lambda:[0.0,0])
for r in rows[1:]:
    v=float(r[iv].replace(",","")); u=r[iu]
    ms=v/1e6 if u in ("nsecond","ns") else v/1e3 if u in ("usecond","us") else v
    agg[r[ik][:110]][0]+=ms; agg[r[ik][:110]][1]+=1
tot=sum(v[0] for v in agg.values())
print("total %.1f ms over %d launches"%(tot,sum(v[1] for v in agg.values())))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][0])[:25]: print("%9.2f ms %5.1f%% n=%5d %s"%(v[0],100*v[0]/tot,v[1],k))
PY
