mkdir -p gpurun_out
cd scripts && timeout 45 python bench_chol_sweep.py > ../gpurun_out/r04g_chol_sweep.jsonl 2> ../gpurun_out/r04g_chol_sweep.err; cd ..
cat gpurun_out/r04g_chol_sweep.jsonl; tail -3 gpurun_out/r04g_chol_sweep.err
