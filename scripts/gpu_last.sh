TAG=${1:-r01z}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_${TAG}_c2_full.json 2> gpurun_out/bench_${TAG}_c2_full.err; tail -c 600 gpurun_out/bench_${TAG}_c2_full.json | head -c 600; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_c3.json 2> gpurun_out/bench_${TAG}_c3.err
python - <<PY
import json
for n in ("c2_full","c3"):
    try:
        d=json.loads(open("gpurun_out/bench_${TAG}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d["roofline"]["kernel"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e)
PY
