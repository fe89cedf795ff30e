TAG=${1:-pool}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -x -q -m gpu 2>&1 | tail -15
for pm in 150000 0; do
APB_POOL_MIN=$pm timeout 900 python bench.py --workload c3s --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3s_$pm.json 2> gpurun_out/bench_${TAG}_c3s_$pm.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c3s_$pm.json')); print('c3s pool_min $pm', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k}, d['roofline_all'])"; tail -3 gpurun_out/bench_${TAG}_c3s_$pm.err
done
for pm in 2000000000 0; do
APB_POOL_MIN=$pm timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_${TAG}_c2_$pm.json 2> gpurun_out/bench_${TAG}_c2_$pm.err; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_c2_$pm.json')); print('c2 pool_min $pm', d['ms_per_step'], {k:v for k,v in d['kernel_ms'].items() if 'integ' in k})"; tail -3 gpurun_out/bench_${TAG}_c2_$pm.err
done
timeout 1500 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err; tail -c 2500 gpurun_out/bench_${TAG}_c3.json; tail -5 gpurun_out/bench_${TAG}_c3.err
