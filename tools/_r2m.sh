V=$PWD/fri-reinforcementlearning-c_b200/build/variants
echo "== isolated"; FRIQ_LIB=$V/friq_tl32.so AGENTS=4736 python tools/timeline.py r2m_iso 20
echo "== loaded"; FRIQ_LIB=$V/friq_tl.so AGENTS=65536 python tools/timeline.py r2m_loaded 20
echo "== loaded rs4"; FRIQ_LIB=$V/friq_tlrs4.so AGENTS=65536 python tools/timeline.py r2m_rs4 20
