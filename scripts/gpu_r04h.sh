mkdir -p gpurun_out
timeout 25 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "cholesky" 2>&1 | tail -3
timeout 20 python scripts/bench_chol.py > gpurun_out/r04h_chol.jsonl 2> gpurun_out/r04h_chol.err; cut -c1-260 gpurun_out/r04h_chol.jsonl
