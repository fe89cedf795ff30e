mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02r_bench.json").read().strip().splitlines()[-1])
print("c3", round(d["value"],3), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "cpu", d["cpu_baseline"]["value"])
print({k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernel_ms"].items())[:12]})
print(d["roofline_all"])
for k,v in d.get("other_workloads",{}).items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("e2e",{}).get("value"), v.get("error"))
    print("   ", {kk:round(vv["ms"]/20,3) for kk,vv in list(v.get("kernel_ms",{}).items())[:10]})
PY
