mkdir -p gpurun_out
for i in $(seq 1 30); do
  CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "large_psf or refinement_queue" > gpurun_out/r02p_run.log 2>&1
  if grep -q "failed" gpurun_out/r02p_run.log; then echo "FAILED at run $i"; grep -n "Error\|error\|apb_\|fit.py\|cabi.py" gpurun_out/r02p_run.log | head -30; cp gpurun_out/r02p_run.log gpurun_out/r02p_fail.log; break; fi
done
echo "loop done"; tail -2 gpurun_out/r02p_run.log
for i in $(seq 1 15); do
  timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "large_psf or refinement_queue" > gpurun_out/r02p_run2.log 2>&1
  if grep -q "failed" gpurun_out/r02p_run2.log; then echo "FAILED (async) at run $i"; grep -n "Error\|error\|apb_\|fit.py\|cabi.py" gpurun_out/r02p_run2.log | head -30; cp gpurun_out/r02p_run2.log gpurun_out/r02p_fail2.log; break; fi
done
echo "loop2 done"
