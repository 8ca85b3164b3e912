set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/perf_probe.py 1000000 10
VGB_EMIT_WAVES=2 python tools/perf_probe.py 1000000 10
VGB_EMIT_WAVES=4 python tools/perf_probe.py 1000000 10
VGB_EMIT_WAVES=1000 python tools/perf_probe.py 1000000 10
ncu --set full --clock-control none --import-source on -k regex:k_group_emit -s 2 -c 1 -o gpurun_out/r2_d python tools/perf_probe.py 1000000 1 > gpurun_out/ncu_r2_d.log 2>&1
