#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/run_config.py 22 32 4 2 100 3 | cut -c1-700
python scratch/bench_stages.py 22 32 4 merkle,fri
