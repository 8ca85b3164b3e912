python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err
tail -c 2500 gpurun_out/bench_r2_b.json
