TAG=${1:-q}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    km=d["kernel_ms"]
    print("$name value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3))
    print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(km.items())[:12]})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
}
run c2 --no-cpu
run c4 --workload c4 --steps 20 --warmup 3 --no-cpu
