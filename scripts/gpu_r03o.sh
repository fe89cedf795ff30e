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
for v in ns0 ns40; do
APB_LIB_PATH=$PWD/build/lib_pcg_$v.so timeout 600 python scripts/pcg_trace.py > gpurun_out/r03o_$v.log 2>&1; summ gpurun_out/r03o_$v.log
done
