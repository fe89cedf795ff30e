# ncu --set full of the sampling kernels on the crowded-field scale model (throughput-bound launches)
TAG=${1:-c3s}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-'k_integrate|k_first|k_select'} -s ${NCU_S:-12} -c ${NCU_C:-8} -o gpurun_out/prof_${TAG} -f python bench.py --workload c3s --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
