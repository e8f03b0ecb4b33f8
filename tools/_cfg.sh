python -m pytest tests -x -q -m gpu -k "replay_bit_exact or free_running or big_path or 65536" 2>&1 | tail -2
echo "== synthetic: $(python tools/quick_tput.py synthetic 1184 2 2>&1 | tail -2 | head -1)"
echo "== g2s2: $(python tools/quick_tput.py g2s2 16384 10 2>&1 | tail -2 | head -1)"
echo "== g10s10: $(python tools/quick_tput.py g10s10 16384 10 2>&1 | tail -2 | head -1)"
