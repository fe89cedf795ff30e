# 2 GPUs: tile-sharded LM tests over NCCL, then c3 (2 tiles), c4 (4 bands/GPU), c2 (2 bands) benches
TAG=${1:-r01d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "two_gpus" 2>&1 | tail -5
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > gpurun_out/bench_${TAG}_$2_n2.json 2> gpurun_out/bench_${TAG}_$2_n2.err; tail -1 gpurun_out/bench_${TAG}_$2_n2.json | cut -c1-600; tail -2 gpurun_out/bench_${TAG}_$2_n2.err | cut -c1-300; }
run 29511 c3 --workload c3 --steps 4 --warmup 3 --no-cpu
run 29512 c4 --workload c4 --steps 20 --warmup 3 --no-cpu
run 29513 c2 --steps 50 --warmup 5 --no-cpu
