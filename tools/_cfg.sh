python -m pytest tests -x -q -m gpu -k "big_path or wide or synthetic" 2>&1 | tail -2
echo "== synthetic: $(python tools/quick_tput.py synthetic 1184 2 2>&1 | tail -2 | head -1)"
