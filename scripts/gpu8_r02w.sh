mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 scripts/bench_allreduce.py 2>&1 | tail -1
for v in peer nccl; do
  if [ $v = nccl ]; then export APB_NO_PEER=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload c2 --steps 20 --warmup 5 --no-extras > gpurun_out/r02w_c2_$v.json 2> gpurun_out/r02w_c2_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/r02w_c2_$v.json').read().strip().splitlines()[-1]); print('c2 N=8 $v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
