mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "cholesky" 2>&1 | tail -4
timeout 60 python scripts/bench_chol.py > gpurun_out/r04f_chol.jsonl 2> gpurun_out/r04f_chol.err; cat gpurun_out/r04f_chol.jsonl | cut -c1-300
timeout 200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
