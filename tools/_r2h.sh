V=fri-reinforcementlearning-c_b200/build/variants
run() { # tag lib policy workload steps warm
  r=$(FRIQ_LIB=$2 python bench.py --workload $4 --steps $5 --warmup $6 --no-cpu-baseline --policy $3 2>/dev/null | grep '^{' | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value=%.4g ms=%.3f e2e=%.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))')
  echo "$1 policy=$3 $4: $r"
}
L=$PWD/fri-reinforcementlearning-c_b200/libfriq_b200.so
for p in 3 2; do run main $L $p c3 20 5; done
for v in nowl noqq nocos none; do for p in 3 2; do run $v $PWD/$V/friq_$v.so $p c3 20 5; done; done
for p in 3 2; do run main $L $p acrobot 6 3; done
for v in nowl none; do for p in 3 2; do run $v $PWD/$V/friq_$v.so $p acrobot 6 3; done; done
