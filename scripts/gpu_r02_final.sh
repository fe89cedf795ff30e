# Round-2 closing visit on one B200: parity tests, smoke, the driver's bench lines (default + reference arm), fp32 lines,
# ncu launch list of the default command and full captures of the hot kernels.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 ) 2>&1 | grep -v "^$\|user\|sys"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
show() { python - "$1" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "roof", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    for k,v in (d.get("other_workloads") or {}).items(): print("   ", k, v.get("value"), v.get("ms_per_step"), (v.get("e2e") or {}).get("value"), v.get("error"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
( time timeout 1500 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err ) 2>&1 | grep real; show gpurun_out/${TAG}_bench_n1.json
( time timeout 1500 python bench.py --impl reference > gpurun_out/${TAG}_bench_n1_reference.json 2> gpurun_out/${TAG}_bench_n1_reference.err ) 2>&1 | grep real; show gpurun_out/${TAG}_bench_n1_reference.json
timeout 900 python bench.py --dtype f32 --no-extras --no-cpu > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; show gpurun_out/${TAG}_bench_c3_f32.json
timeout 900 python bench.py --workload c5s --no-extras --no-cpu > gpurun_out/${TAG}_bench_c5s_f64.json 2> gpurun_out/${TAG}_bench_c5s_f64.err; show gpurun_out/${TAG}_bench_c5s_f64.json
timeout 900 python bench.py --workload c5s --dtype f32 --no-extras --no-cpu > gpurun_out/${TAG}_bench_c5s_f32.json 2> gpurun_out/${TAG}_bench_c5s_f32.err; show gpurun_out/${TAG}_bench_c5s_f32.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_pool|k_pcg|k_blocks|k_fft_cols|k_select|k_block_gather' -s 40 -c 24 -o gpurun_out/${TAG}_prof_c3 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof_c3.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_c3_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep ${TAG}_
