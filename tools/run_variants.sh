#!/bin/bash
# run bench.py for every engine variant built by tools/build_variants.sh (on the GPU box); one line per variant / setting
cd "$(dirname "$0")/.."
for so in fri-reinforcementlearning-c_b200/build/variants/friq_*.so; do
  for g in ${GROUPS_LIST:-0}; do
   {
    r=$(FRIQ_LIB=$PWD/$so python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline --groups $g ${BENCH_ARGS} 2>&1 | tail -1)
    echo "$(basename $so) groups=$g $(echo "$r" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.4g ms=%.3f e2e=%.4g frac=%.3f meanR=%.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["stats"]["mean_rules"]))' 2>&1 | tail -1)"
   }
  done
done
