V=fri-reinforcementlearning-c_b200/build/variants
run() { r=$(FRIQ_LIB=$2 python bench.py --workload $4 --steps $5 --warmup $6 --no-cpu-baseline --policy $3 2>/dev/null | grep '^{' | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.4g ms=%.3f e2e=%.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))'); echo "$1 policy=$3 $4: $r"; }
for v in b128 b64 b32; do run $v $PWD/$V/friq_$v.so 2 c3 20 5; done
for v in b128 b32; do run $v $PWD/$V/friq_$v.so 2 acrobot 6 3; done
python -m pytest tests -m gpu -x -q -k "other_action" 2>&1 | tail -3
