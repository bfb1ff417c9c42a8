python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools_perf_probe.py 2>&1 | grep -v "^$"
