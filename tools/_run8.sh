TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $1 --steps 30 --warmup 5 > gpurun_out/b_n$1.json 2> gpurun_out/b_n$1.err; tail -c 1800 gpurun_out/b_n$1.json; tail -2 gpurun_out/b_n$1.err
$TR bench.py --gpus $1 --steps 30 --warmup 5 --merged --gather > gpurun_out/b_n$1_merged.json 2> gpurun_out/b_n$1_merged.err; tail -c 2400 gpurun_out/b_n$1_merged.json; tail -2 gpurun_out/b_n$1_merged.err
$TR bench.py --gpus $1 --workload tiger_x10000 --steps 5 --warmup 3 --gather --e2e-steps 1 > gpurun_out/b_n$1_tiger10k.json 2> gpurun_out/b_n$1_tiger10k.err; tail -c 3000 gpurun_out/b_n$1_tiger10k.json; tail -3 gpurun_out/b_n$1_tiger10k.err
