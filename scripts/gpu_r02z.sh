mkdir -p gpurun_out
timeout 600 python scripts/pcg_trace.py > gpurun_out/r02z_pcg_trace.log 2>&1
head -40 gpurun_out/r02z_pcg_trace.log
