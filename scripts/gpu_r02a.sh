# Round 2, first visit: the 18 never-run tests, then baselines of c3 / c4 / c2 with the round-1 code
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
APB_ALLOW_UNVERIFIED=1 timeout 900 python -m pytest tests/test_cuda_unverified.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r02a_unverified.log; tail -8 gpurun_out/r02a_unverified.log
( time timeout 900 python bench.py --workload c3 --steps 4 --warmup 3 > gpurun_out/r02a_c3.json 2> gpurun_out/r02a_c3.err ) 2>&1 | grep real
( time timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r02a_c4.json 2> gpurun_out/r02a_c4.err ) 2>&1 | grep real
( time timeout 600 python bench.py --workload c3 --impl reference --steps 2 --warmup 1 > gpurun_out/r02a_c3ref.json 2> gpurun_out/r02a_c3ref.err ) 2>&1 | grep real
python - <<'PY'
import json
for n in ("c3","c4","c3ref"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02a_{n}.err").read()[-1500:])
PY
