#!/bin/bash
# The end-of-round measurement on ONE B200 (run through gpurun): parity tests, smoke, every bench workload with its CPU baseline,
# the reference arm, the calibration captures (tools/calibrate.sh) and the launch list.  Outputs: gpurun_out/${TAG}_*.
TAG=${TAG:-r2b}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
bash tools/calibrate.sh > gpurun_out/${TAG}_calibrate.log 2>&1; tail -3 gpurun_out/${TAG}_calibrate.log
python tools/calibrate.py fit >> gpurun_out/${TAG}_calibrate.log 2>&1; tail -2 gpurun_out/${TAG}_calibrate.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3_reference.json 2> gpurun_out/${TAG}_bench_c3_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
for w in acrobot cartpole c4 c5 c1 c2; do
  python bench.py --workload $w --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err
done
for w in c3_reference c3 acrobot cartpole c4 c5 c1 c2; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$w.json"))
    ex = (d.get("roofline") or {}).get("executed") or {}
    print("$w", "%.4g" % d["value"], "ms %.3f" % (d.get("ms_per_step") or 0), "e2e %.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"),
          "frac_executed", ex.get("frac_executed"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$w", "FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
tail -1 gpurun_out/${TAG}_launches.csv | cut -c1-200
