mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r02i_c4.json 2> gpurun_out/r02i_c4.err
timeout 600 python bench.py --workload c2 --steps 50 --warmup 3 --no-cpu > gpurun_out/r02i_c2.json 2> gpurun_out/r02i_c2.err
timeout 600 python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02i_c3.json 2> gpurun_out/r02i_c3.err
python - <<'PY'
import json
for n in ("c4","c2","c3"):
    try:
        d=json.loads(open(f"gpurun_out/r02i_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))
        print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(d["kernel_ms"].items())[:9]})
        print("   ", {k:v for k,v in d["roofline_all"].items() if "fft" in k})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02i_{n}.err").read()[-1500:])
PY
