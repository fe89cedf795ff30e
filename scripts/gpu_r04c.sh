mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -25
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/r04c_bench.json 2> gpurun_out/r04c_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r04c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:round(v['ms']/d['steps'],2) for k,v in list(d['kernel_ms'].items())[:8]})
PY
