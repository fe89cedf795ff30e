# Round 1, visit d: tile-sharding tests, default bench with cpu_baseline, c3 / c4 on one GPU
TAG=${1:-r01d}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; tail -c 800 gpurun_out/bench_${TAG}_ref.json
timeout 900 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_c4.json 2> gpurun_out/bench_${TAG}_c4.err; tail -c 1500 gpurun_out/bench_${TAG}_c4.json; tail -3 gpurun_out/bench_${TAG}_c4.err
timeout 1200 python bench.py --workload c3 --steps 4 --warmup 3 > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err; tail -c 1500 gpurun_out/bench_${TAG}_c3.json; tail -3 gpurun_out/bench_${TAG}_c3.err
