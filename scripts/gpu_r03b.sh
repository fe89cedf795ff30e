mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python scripts/pcg_trace.py > gpurun_out/r03b_$tag.log 2>&1; python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/r03b_$tag.log").read().strip().splitlines()[1:]]
its=sum(r["its"][0] for r in rows); ms=sum(r["ms"] for r in rows)
print("$tag: solves", len(rows), "iterations", its, "ms", round(ms,2), "us/iteration", round(1e3*ms/its,2))
PY
}
run sleep200 APB_X=1
run sleep200_1persm APB_PCG_PER_SM=1
run spin0_1persm APB_PCG_PER_SM=1 APB_LIB_PATH=$PWD/build/lib_spin0.so
