python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/_dbg.py 300
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'])"
python bench.py --steps 100 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['stage_ms'])"
