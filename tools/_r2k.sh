V=$PWD/fri-reinforcementlearning-c_b200/build/variants
python -m pytest tests -m gpu -x -q -k "policy or table_mode or exploration or bench_workload or golden or dictionary" 2>&1 | tail -3
run() { r=$(FRIQ_LIB=$2 python bench.py --workload $4 --steps $5 --warmup $6 --no-cpu-baseline --policy $3 2>/dev/null | grep '^{' | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.4g ms=%.3f e2e=%.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))'); echo "$1 policy=$3 $4: $r"; }
run rs4 $PWD/fri-reinforcementlearning-c_b200/libfriq_b200.so 2 c3 20 5
for v in rs0 rs2 rs8; do run $v $V/friq_$v.so 2 c3 20 5; done
run rs4 $PWD/fri-reinforcementlearning-c_b200/libfriq_b200.so 2 acrobot 6 3
run rs4 $PWD/fri-reinforcementlearning-c_b200/libfriq_b200.so 2 c5 6 3
echo "== timeline rs4 loaded"; FRIQ_LIB=$V/friq_tlrs4.so AGENTS=65536 python tools/timeline.py r2_rs4 20
