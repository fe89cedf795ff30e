# crowded-field bring-up on one B200: solver parity tests, bench on the scale model and on full config[2]
TAG=${1:-c3}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "${TESTK:-solver or pcg or crowded}" 2>&1 | tail -15
timeout 900 python bench.py --workload c3s --steps ${S1:-10} --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3s.json 2> gpurun_out/bench_${TAG}_c3s.err; tail -c 3000 gpurun_out/bench_${TAG}_c3s.json; tail -5 gpurun_out/bench_${TAG}_c3s.err
timeout 1500 python bench.py --workload c3 --steps ${S2:-4} --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err; tail -c 3000 gpurun_out/bench_${TAG}_c3.json; tail -5 gpurun_out/bench_${TAG}_c3.err
