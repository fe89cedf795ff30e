# Round 2, visit b: table-driven exp/log in the profile kernels -- parity, then c3 / c4 / c2 / c5s timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for v in default minb4; do
  if [ $v = minb4 ]; then export APB_LIB_PATH=$PWD/build/lib_minb4.so; fi
  timeout 600 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r02b_c3_$v.json 2> gpurun_out/r02b_c3_$v.err
done
unset APB_LIB_PATH
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02b_c4.json 2> gpurun_out/r02b_c4.err
timeout 600 python bench.py --workload c2 --steps 50 --warmup 3 --no-cpu > gpurun_out/r02b_c2.json 2> gpurun_out/r02b_c2.err
timeout 600 python bench.py --workload c5s --steps 10 --warmup 3 --no-cpu > gpurun_out/r02b_c5s.json 2> gpurun_out/r02b_c5s.err
python - <<'PY'
import json
for n in ("c3_default","c3_minb4","c4","c2","c5s"):
    try:
        d=json.loads(open(f"gpurun_out/r02b_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))
        print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(d["kernel_ms"].items())[:9]})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02b_{n}.err").read()[-1500:])
PY
