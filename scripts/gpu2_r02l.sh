mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "two_gpus" 2>&1 | tail -15
for wl in c4 c3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 > gpurun_out/r02l_${wl}_n2.json 2> gpurun_out/r02l_${wl}_n2.err
APB_NO_PEER=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 > gpurun_out/r02l_${wl}_n2_nccl.json 2> gpurun_out/r02l_${wl}_n2_nccl.err
done
python - <<'PY'
import json
for n in ("c4_n2","c4_n2_nccl","c3_n2","c3_n2_nccl"):
    try:
        d=json.loads(open(f"gpurun_out/r02l_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r02l_{n}.err").read()[-1500:])
PY
