mkdir -p gpurun_out
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r02f_bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02f_ref.json 2> gpurun_out/r02f_ref.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r02f_ref.json").read().strip().splitlines()[-1])
print("c3", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"], "ref arm", r["value"], r["steps_timed"])
print(d["roofline"]); print(d["roofline_all"]); print(d["config"]["timing"])
for k,v in d.get("other_workloads",{}).items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("e2e",{}).get("value"), v.get("error"), v.get("roofline_all"))
PY
