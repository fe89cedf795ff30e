TAG=${1:-r01w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -m gpu -k "two_gpus" 2>&1 | tail -3
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > gpurun_out/bench_${TAG}_$2_n2.json 2> gpurun_out/bench_${TAG}_$2_n2.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$2_n2.json").read().strip().splitlines()[-1])
    print("$2 n=2 value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d.get("impl"), d["scaling"])
except Exception as e:
    print("$2 FAILED", e); print(open("gpurun_out/bench_${TAG}_$2_n2.err").read()[-2500:])
PY
}
run 29611 c2 --steps 50 --warmup 5
run 29612 ref --impl reference --steps 2 --warmup 1
run 29613 c3s --workload c3s --steps 10 --warmup 3 --no-cpu
run 29614 c4 --workload c4 --steps 10 --warmup 3 --no-cpu
