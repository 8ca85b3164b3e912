python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15
python tools/perf_probe_tiger.py 600 2>&1 | tail -2
