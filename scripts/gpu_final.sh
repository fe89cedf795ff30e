# Round-1 closing visit: parity tests, smoke, bench lines (default + reference arm + other workloads), ncu launch list + full capture
TAG=${1:-r01m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
run() { name=$1; shift; timeout 1200 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    km=d.get("kernel_ms",{})
    print("$name value", round(d["value"],3), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), "P", d["config"].get("params"), "pcg", d["config"].get("pcg_iterations_mean"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "roof", d.get("roofline",{}).get("kernel"), d.get("roofline",{}).get("frac"))
    print("   ", {k:round(v["ms"]/d["steps"],3) for k,v in list(km.items())[:8]})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
}
run c2
run ref --impl reference --steps 5 --warmup 1
run c3 --workload c3 --steps 4 --warmup 3
run c4 --workload c4 --steps 20 --warmup 3
run c5s --workload c5s --steps 10 --warmup 3
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_cols|k_fft_rows|k_blocks|k_integrate|k_first|k_geo_v' -s 60 -c 16 -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'k_integrate_pool' -s 4 -c 2 -o gpurun_out/prof_${TAG}_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${TAG}_c3.log 2>&1
ncu -i gpurun_out/prof_${TAG}_c3.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_c3_raw.csv 2>/dev/null
for f in gpurun_out/prof_${TAG}.ncu-rep gpurun_out/prof_${TAG}_c3.ncu-rep; do if [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ]; then rm -f $f; fi; done
ls -la gpurun_out/ | grep ${TAG}
fi
