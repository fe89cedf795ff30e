TAG=${1:-pcg}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 20 -c 1 -o gpurun_out/prof_${TAG} -f python bench.py --workload ${WL:-c3s} --steps 2 --warmup 2 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
