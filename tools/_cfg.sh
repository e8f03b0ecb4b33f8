python -m pytest tests -x -q -m gpu -k "big or wide or synthetic" 2>&1 | tail -3
echo "== synthetic: $(python tools/quick_tput.py synthetic 1184 2 2>&1 | tail -2 | head -1)"
echo "== synthetic 512x2: $(SER_BIG_THREADS=512 SER_BIG_SMEM_KB=110 python tools/quick_tput.py synthetic 1184 2 2>&1 | tail -2 | head -1)"
