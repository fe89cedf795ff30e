# usage: bash scripts/gpuN_scale.sh N TAG   -- the driver's invocation at N GPUs (default workload + extras)
N=$1; TAG=$2
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err ) 2>&1 | grep real
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
    print("c3 N=$N", round(d["value"],3), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), d["config"]["pcg_iterations_mean"])
    print({k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernel_ms"].items())[:8]})
    for k,v in d.get("other_workloads",{}).items():
        print(k, v.get("value"), v.get("ms_per_step"), v.get("e2e",{}).get("value"), v.get("error"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${TAG}_n$N.err").read()[-2500:])
PY
