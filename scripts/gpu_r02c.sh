mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "test_refinement_queue_grows_on_overflow" 2>&1 | grep -v "^$" | head -60 > gpurun_out/r02c_sanitizer.log
tail -40 gpurun_out/r02c_sanitizer.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_pool' -s 6 -c 1 -o gpurun_out/prof_r02c_pool -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
ncu -i gpurun_out/prof_r02c_pool.ncu-rep --page raw --csv > gpurun_out/prof_r02c_pool_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02c_pool.ncu-rep --page source --csv > gpurun_out/prof_r02c_pool_source.csv 2>/dev/null
ls -la gpurun_out | grep r02c
