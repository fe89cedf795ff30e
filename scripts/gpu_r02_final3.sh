mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 700 -c 460 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
tail -3 gpurun_out/r02_ncu_launches.log | cut -c1-300
wc -l gpurun_out/r02_launches_c3.csv
