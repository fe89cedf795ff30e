# ncu --set full of one k_integrate_pool launch (value-only pass) at config[2] with the pooled depth-3 phase
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_pool' -s 6 -c 1 -o gpurun_out/prof_r04a_pool -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/r04a_ncu.log 2>&1
tail -3 gpurun_out/r04a_ncu.log | cut -c1-300
ncu -i gpurun_out/prof_r04a_pool.ncu-rep --page raw --csv > gpurun_out/prof_r04a_pool_raw.csv 2>/dev/null
ls -la gpurun_out/prof_r04a*
