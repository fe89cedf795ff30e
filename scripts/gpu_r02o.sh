mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "fft_convolution or refinement_queue" 2>&1 | tail -3
done
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "fft_convolution or refinement_queue" 2>&1 | grep -v "^$" | head -80 > gpurun_out/r02o_memcheck.log
tail -40 gpurun_out/r02o_memcheck.log
