V=$PWD/fri-reinforcementlearning-c_b200/build/variants
echo "== loaded: 65536 agents, 128-thread CTAs"; FRIQ_LIB=$V/friq_tl.so AGENTS=65536 python tools/timeline.py r2_loaded 20
echo "== isolated: 4736 agents = 148 tiles, 32-thread CTAs (one warp per SM)"; FRIQ_LIB=$V/friq_tl32.so AGENTS=4736 python tools/timeline.py r2_iso 20
echo "== half: 18944 agents = 592 tiles, 32-thread CTAs (4 warps per SM)"; FRIQ_LIB=$V/friq_tl32.so AGENTS=18944 python tools/timeline.py r2_half 20
echo "== 37888 agents = 1184 tiles, 128-thread CTAs (one wave)"; FRIQ_LIB=$V/friq_tl.so AGENTS=37888 python tools/timeline.py r2_wave 20
python -m pytest tests -m gpu -x -q -k "other_action or policy or table_mode" 2>&1 | tail -3
