mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -k "fp32 or float32 or eight_band or pooled or fused_integration" 2>&1 | tail -30
for dt in f64 f32; do
timeout 600 python bench.py --dtype $dt --workload c5s --steps 10 --warmup 3 --no-cpu > gpurun_out/r02h_c5s_$dt.json 2> gpurun_out/r02h_c5s_$dt.err
timeout 600 python bench.py --dtype $dt --workload c3 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02h_c3_$dt.json 2> gpurun_out/r02h_c3_$dt.err
done
python - <<'PY'
import json
for n in ("c5s_f64","c5s_f32","c3_f64","c3_f32"):
    try:
        d=json.loads(open(f"gpurun_out/r02h_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))
        print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(d["kernel_ms"].items())[:9]})
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02h_{n}.err").read()[-1500:])
PY
