mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x -k "two_gpu or peer or nccl or distributed or tile" 2>&1 | tail -3
