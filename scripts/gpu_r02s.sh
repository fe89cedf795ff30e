mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x -k "normal_equations or lm_fit or fullsize or tiled" 2>&1 | tail -4
for wl in c4 c2 c3; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > gpurun_out/r02s_$wl.json 2> gpurun_out/r02s_$wl.err
done
python - <<'PY'
import json
for n in ("c4","c2","c3"):
    try:
        d=json.loads(open(f"gpurun_out/r02s_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernel_ms"].items() if k in ("k_blocks","k_geo_v","k_assemble","k_select")}, d["roofline_all"].get("k_blocks"))
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02s_{n}.err").read()[-1500:])
PY
