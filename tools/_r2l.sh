V=$PWD/fri-reinforcementlearning-c_b200/build/variants
for v in tlbase tlnoq tlnotab tlnone; do echo "== isolated $v"; FRIQ_LIB=$V/friq_$v.so AGENTS=4736 python tools/timeline.py r2_$v 20; done
python -m pytest tests/test_dropin_gpu.py -x -q -k "mpi" 2>&1 | tail -3
