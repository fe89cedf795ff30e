TAG=${1:-r01p}
N=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29888 bench.py --gpus $N --workload c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/scale_${TAG}_c3_n$N.json 2> gpurun_out/scale_${TAG}_c3_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${TAG}_c3_n$N.json").read().strip().splitlines()[-1])
    km=d["kernel_ms"]
    print("c3 n=$N value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "pcg", d["config"].get("pcg_iterations_mean"))
    print("   ", {k:round(v["ms"]/d["steps"],2) for k,v in list(km.items())[:6]})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/scale_${TAG}_c3_n$N.err").read()[-2000:])
PY
