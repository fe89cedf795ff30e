mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras > gpurun_out/r03p_bench.json 2> gpurun_out/r03p_bench.err; tail -c 3000 gpurun_out/r03p_bench.json
