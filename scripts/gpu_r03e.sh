mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "pcg or sparse or large_system or crowded or tiled" 2>&1 | tail -3
timeout 600 python scripts/pcg_trace.py > gpurun_out/r03e.log 2>&1; python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/r03e.log").read().strip().splitlines()[1:]]
its=sum(r["its"][0] for r in rows); ms=sum(r["ms"] for r in rows)
print("solves", len(rows), "iterations", its, "ms", round(ms,2), "us/iteration", round(1e3*ms/its,2))
PY
