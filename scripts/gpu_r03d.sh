mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "pcg or sparse or large_system or crowded or tiled" 2>&1 | tail -3
run() { tag=$1; shift; env "$@" timeout 600 python scripts/pcg_trace.py > gpurun_out/r03d_$tag.log 2>&1; python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/r03d_$tag.log").read().strip().splitlines()[1:]]
its=sum(r["its"][0] for r in rows); ms=sum(r["ms"] for r in rows)
print("$tag: solves", len(rows), "iterations", its, "ms", round(ms,2), "us/iteration", round(1e3*ms/its,2))
PY
}
run regs172 APB_X=1
run regs128 APB_LIB_PATH=$PWD/build/lib_pcg2.so
