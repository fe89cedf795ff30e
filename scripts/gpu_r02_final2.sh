# Round-2 closing visit, part 2 (bench lines + short ncu passes; large reports are reduced to CSV and deleted on the box)
TAG=${1:-r02}
mkdir -p gpurun_out
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
timeout 1500 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; show gpurun_out/${TAG}_bench_n1.json
timeout 1500 python bench.py --impl reference > gpurun_out/${TAG}_bench_n1_reference.json 2> gpurun_out/${TAG}_bench_n1_reference.err; show gpurun_out/${TAG}_bench_n1_reference.json
timeout 900 python bench.py --dtype f32 --no-extras --no-cpu > gpurun_out/${TAG}_bench_c3_f32.json 2> gpurun_out/${TAG}_bench_c3_f32.err; show gpurun_out/${TAG}_bench_c3_f32.json
timeout 900 python bench.py --workload c5s --no-extras --no-cpu > gpurun_out/${TAG}_bench_c5s_f64.json 2> gpurun_out/${TAG}_bench_c5s_f64.err; show gpurun_out/${TAG}_bench_c5s_f64.json
timeout 900 python bench.py --workload c5s --dtype f32 --no-extras --no-cpu > gpurun_out/${TAG}_bench_c5s_f32.json 2> gpurun_out/${TAG}_bench_c5s_f32.err; show gpurun_out/${TAG}_bench_c5s_f32.json
timeout 900 python bench.py --workload c2 --no-extras --no-cpu > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; show gpurun_out/${TAG}_bench_c2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 450 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pcg$' -s 9 -c 1 -o gpurun_out/${TAG}_prof_pcg -f python scripts/pcg_trace.py > gpurun_out/${TAG}_ncu_pcg.log 2>&1
ncu -i gpurun_out/${TAG}_prof_pcg.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_pcg_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof_pcg.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_prof_pcg_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
