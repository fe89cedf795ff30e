mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r03q_bench.json 2> gpurun_out/r03q_bench.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r03q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:round(v['ms'],1) for k,v in d['kernel_ms'].items()})
PY
