TAG=${1:-pool2}
mkdir -p gpurun_out
for v in default 3 5 6; do
  if [ $v = default ]; then unset APB_LIB_PATH; else export APB_LIB_PATH=$PWD/build/variants/lib_mb$v.so; fi
  timeout 900 python bench.py --workload c3s --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3s_$v.json 2> gpurun_out/bench_${TAG}_c3s_$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c3s_$v.json')); print('c3s minb $v', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k}, d['roofline']['frac'])"; tail -3 gpurun_out/bench_${TAG}_c3s_$v.err
done
unset APB_LIB_PATH
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_pool' -s 4 -c 3 -o gpurun_out/prof_${TAG} -f python bench.py --workload c3s --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/prof_${TAG}*
