mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 60 python scripts/bench_chol.py > gpurun_out/r04e_chol.jsonl 2> gpurun_out/r04e_chol.err; cat gpurun_out/r04e_chol.jsonl | cut -c1-300
timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cuda_parity.py -q -x -k "supersampled_psf_both or (cholesky and (33 or 333)) or point_psf_model_up2" > gpurun_out/r04e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r04e_memcheck.log | cut -c1-300
