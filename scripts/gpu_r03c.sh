mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pcg' -s 10 -c 1 -o gpurun_out/prof_r03c_pcg -f python scripts/pcg_trace.py > gpurun_out/r03c_ncu.log 2>&1
ncu -i gpurun_out/prof_r03c_pcg.ncu-rep --page raw --csv > gpurun_out/prof_r03c_pcg_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r03c_pcg.ncu-rep --page source --csv > gpurun_out/prof_r03c_pcg_source.csv 2>/dev/null
ls -la gpurun_out | grep r03c
