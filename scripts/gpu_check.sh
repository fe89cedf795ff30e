# One GPU-box visit: parity tests, bench (both convolution families), ncu launch list + full capture.
# Usage (from the repo root on the box): [NO_NCU=1] [NCU_K=regex] bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python bench.py --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 4000 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 300 python bench.py --no-cpu --conv direct --steps 20 > gpurun_out/bench_${TAG}_direct.json 2>/dev/null; tail -c 600 gpurun_out/bench_${TAG}_direct.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:${NCU_K:-'k_fft|k_blocks|k_refine|k_first'} -s ${NCU_S:-60} -c ${NCU_C:-16} -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
# gpurun_out/ travels back only below 64 MiB: keep the report if it is small, the CSV always
if [ $(stat -c %s gpurun_out/prof_${TAG}.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/prof_${TAG}.ncu-rep; fi
ls -la gpurun_out/
fi
