TAG=${1:-r01q}
mkdir -p gpurun_out
timeout 1500 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c5.json 2> gpurun_out/bench_${TAG}_c5.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_c5.json").read().strip().splitlines()[-1])
    km=d["kernel_ms"]
    print("c5 value", round(d["value"],4), "ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],4), "P", d["config"]["params"], "pcg", d["config"].get("pcg_iterations_mean"), d["config"].get("block_array_doubles"))
    print("   ", {k:round(v["ms"]/d["steps"],1) for k,v in list(km.items())[:8]}, d["refine_queue_last"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_${TAG}_c5.err").read()[-2000:])
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
