mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r02e_c3.json 2> gpurun_out/r02e_c3.err
timeout 600 python bench.py --workload c2 --steps 50 --warmup 3 --no-cpu > gpurun_out/r02e_c2.json 2> gpurun_out/r02e_c2.err
timeout 600 python bench.py --workload c5s --steps 10 --warmup 3 --no-cpu > gpurun_out/r02e_c5s.json 2> gpurun_out/r02e_c5s.err
python - <<'PY'
import json
for n in ("c3","c2","c5s"):
    try:
        d=json.loads(open(f"gpurun_out/r02e_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))
        print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(d["kernel_ms"].items())[:9]})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02e_{n}.err").read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_pool' -s 6 -c 1 -o gpurun_out/prof_r02e_pool -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02e_ncu.log 2>&1
ncu -i gpurun_out/prof_r02e_pool.ncu-rep --page raw --csv > gpurun_out/prof_r02e_pool_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r02e_pool.ncu-rep --page source --csv > gpurun_out/prof_r02e_pool_source.csv 2>/dev/null
