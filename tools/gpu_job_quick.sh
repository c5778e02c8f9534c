timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/prove_once.py 22 32 4 4 | tail -2 | cut -c1-900
python tools/prove_once.py 20 16 8 3 | tail -1 | cut -c1-700
