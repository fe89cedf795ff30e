mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for v in default minb6; do
  if [ $v != default ]; then export APB_LIB_PATH=$PWD/build/lib_$v.so; fi
  timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02y_c3_$v.json 2> gpurun_out/r02y_c3_$v.err
done
unset APB_LIB_PATH
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02y_c2.json 2> gpurun_out/r02y_c2.err
python - <<'PY'
import json
for n in ("c3_default","c3_minb6","c2"):
    try:
        d=json.loads(open(f"gpurun_out/r02y_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "trials", d["config"]["lambda_trials_per_iter"], "restarts", d["config"]["fit_restarts"], "blocks", len(d["blocks_ms"]))
        print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(d["kernel_ms"].items())[:8]})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02y_{n}.err").read()[-1500:])
PY
