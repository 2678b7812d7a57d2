#!/bin/bash
# round 2, call H: ncu of the tcgen05 prefill mat-mul + launch list of one 256-token batch (2 layers)
mkdir -p gpurun_out
timeout 300 python tools/prompt_probe.py --layers 2 --n 256 > gpurun_out/r2h_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python tools/prompt_probe.py --layers 2 --n 256 --reps 1 > gpurun_out/r2h_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:q4_gemm_tc -s 1 -c 2 -f -o gpurun_out/r2h_tc python tools/prompt_probe.py --layers 2 --n 256 --reps 1 > gpurun_out/r2h_ncu_full.log 2>&1
cat gpurun_out/r2h_probe.log; tail -3 gpurun_out/r2h_ncu_full.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2h_launches.csv")) if len(r) > 5]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    agg[r[kn].split("(")[0]][0] += 1; agg[r[kn].split("(")[0]][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{k:50s} {v[0]:5d} launches {v[1]/1e3:10.1f} us {100*v[1]/tot:5.1f}%")
PY
