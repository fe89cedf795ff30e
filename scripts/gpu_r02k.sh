mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_cols|k_fft_rows' -s 40 -c 6 -o gpurun_out/prof_r02k_fft -f python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02k_ncu.log 2>&1
ncu -i gpurun_out/prof_r02k_fft.ncu-rep --page raw --csv > gpurun_out/prof_r02k_fft_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02k_fft.ncu-rep --page source --csv > gpurun_out/prof_r02k_fft_source.csv 2>/dev/null
ls -la gpurun_out | grep r02k
