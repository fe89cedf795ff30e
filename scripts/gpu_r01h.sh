TAG=${1:-r01h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    km=d["kernel_ms"]
    print("$name value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "P", d["config"]["params"], "trials", d["config"]["lambda_trials_per_iter"], "pcg", d["config"].get("pcg_iterations_mean"), d["config"].get("pcg_solves"), d["config"].get("block_array_doubles"))
    print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(km.items())[:10]}, d["refine_queue_last"], d["roofline"].get("frac"))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
}
run c3s --workload c3s --steps 10 --warmup 3 --no-cpu
run c3 --workload c3 --steps 4 --warmup 3 --no-cpu
