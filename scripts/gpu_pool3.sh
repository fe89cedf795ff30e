TAG=${1:-pool3}
mkdir -p gpurun_out
for fl in 0 4; do
APB_PLAN_FLAGS=$fl timeout 900 python bench.py --workload c3s --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3s_$fl.json 2> gpurun_out/bench_${TAG}_c3s_$fl.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c3s_$fl.json')); print('c3s flags $fl', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k or 'refine' in k or 'reduce' in k or 'scatter' in k})"; tail -3 gpurun_out/bench_${TAG}_c3s_$fl.err
done
for v in default 3 5 6; do
  if [ $v = default ]; then unset APB_LIB_PATH; else export APB_LIB_PATH=$PWD/build/variants/lib_mb$v.so; fi
  timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3_$v.json 2> gpurun_out/bench_${TAG}_c3_$v.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c3_$v.json')); print('c3 minb $v', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k}, d['roofline']['frac'])"; tail -3 gpurun_out/bench_${TAG}_c3_$v.err
done
unset APB_LIB_PATH
APB_PLAN_FLAGS=4 timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3_f4.json 2> gpurun_out/bench_${TAG}_c3_f4.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c3_f4.json')); print('c3 flags 4', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k or 'refine' in k or 'reduce' in k or 'scatter' in k})"; tail -3 gpurun_out/bench_${TAG}_c3_f4.err
