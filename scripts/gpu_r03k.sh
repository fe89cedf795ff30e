mkdir -p gpurun_out
summ() { python - "$1" <<PY
import json,sys
rows=[]
for l in open(sys.argv[1]).read().strip().splitlines()[1:]:
    if l.startswith("{"):
        r=json.loads(l)
        if "its" in r: rows.append(r)
    else: print(l)
its=sum(r["its"][0] for r in rows); ms=sum(r["ms"] for r in rows)
print(sys.argv[1], "solves", len(rows), "iterations", its, "ms", round(ms,2), "us/iteration", round(1e3*ms/its,2))
PY
}
APB_PCG_PER_SM=1 APB_LIB_PATH=$PWD/build/lib_pcg1.so timeout 600 python scripts/pcg_trace.py > gpurun_out/r03k_minb1.log 2>&1; summ gpurun_out/r03k_minb1.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pcg$ -s 9 -c 1 -o gpurun_out/r03k_pcg -f python scripts/pcg_trace.py > gpurun_out/r03k_ncu.log 2>&1; tail -2 gpurun_out/r03k_ncu.log
