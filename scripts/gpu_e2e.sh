TAG=${1:-r01n}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    print("$name value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "P", d["config"]["params"], "restarts", d["config"]["fit_restarts"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-2500:])
PY
}
run c2 --no-cpu
run c2b --workload c2b --no-cpu
run c3s --workload c3s --steps 10 --warmup 3 --no-cpu
run c4 --workload c4 --steps 20 --warmup 3 --no-cpu
