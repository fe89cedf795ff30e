mkdir -p gpurun_out
fails=0
for i in $(seq 1 40); do
  timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "refinement_queue" > gpurun_out/r02q_run.log 2>&1
  if grep -q "failed" gpurun_out/r02q_run.log; then echo "FAILED at run $i"; fails=1; grep -n "Error\|error" gpurun_out/r02q_run.log | head -10; break; fi
done
echo "overflow loop done, fails=$fails"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6
