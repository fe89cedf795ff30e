mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25
