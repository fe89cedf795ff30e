mkdir -p gpurun_out
run() { tag=$1; wl=$2; shift 2; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload $wl --steps 20 --warmup 5 --no-extras > gpurun_out/r02u_$tag.json 2> gpurun_out/r02u_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02u_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["value"],3), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), {k:round(v["ms"]/d["steps"],2) for k,v in list(d["kernel_ms"].items())[:4]})
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r02u_$tag.err").read()[-1200:])
PY
}
run c2_peer c2 APB_X=1
run c2_nccl c2 APB_NO_PEER=1
run c3_2x4 c3 APB_X=1
run c3_4x4 c3 APB_TILES=4x4
run c3_4x8 c3 APB_TILES=4x8
run c3_2x4_nccl c3 APB_NO_PEER=1
