ncu --set full --clock-control none --import-source on -k regex:"k_group_emit|k_group_count|k_flatten_slots|k_job_prep" -s 8 -c 4 -o gpurun_out/r02_final python tools/perf_probe.py 1000000 1 > gpurun_out/ncu_r02_final.log 2>&1
tail -3 gpurun_out/ncu_r02_final.log
