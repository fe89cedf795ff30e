# 8-GPU box: scaling of the default bench (c2, weak), c4 (8 bands, strong) and c3 (tiles, strong)
TAG=${1:-r01e}
mkdir -p gpurun_out
run() { # port N name args...
  local port=$1 n=$2 name=$3; shift 3
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_${TAG}_${name}_n$n.json 2> gpurun_out/scale_${TAG}_${name}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/scale_${TAG}_${name}_n$n.json 2> gpurun_out/scale_${TAG}_${name}_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${TAG}_${name}_n$n.json").read().strip().splitlines()[-1])
    print("${name} n=$n value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d["scaling"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name} n=$n FAILED", e); print(open("gpurun_out/scale_${TAG}_${name}_n$n.err").read()[-1500:])
PY
}
for n in 1 2 4 8; do run $((29600+n)) $n c2 --steps 50 --warmup 5 --no-cpu; done
for n in 4 8; do run $((29700+n)) $n c4 --workload c4 --steps 20 --warmup 3 --no-cpu; done
for n in 4 8; do run $((29800+n)) $n c3 --workload c3 --steps 4 --warmup 3 --no-cpu; done
