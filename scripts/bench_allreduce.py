"""Latency of apb_allreduce (NVLink peer memory) against the NCCL all-reduce, back to back on one stream.
torchrun --nproc-per-node N scripts/bench_allreduce.py"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from astrophot_b200.cabi import PeerComm
comm = PeerComm(1 << 20)
out = {}
for n in (64, 212, 22004, 513711):
    t = torch.ones(n, dtype=torch.float64, device="cuda")
    for name, fn in (("peer", lambda: comm.allreduce(t)), ("nccl", lambda: dist.all_reduce(t))):
        for _ in range(20):
            fn(); t.fill_(1.0)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record(); torch.cuda.synchronize()
        out[f"{name}_{n}"] = round(e0.elapsed_time(e1) / 200 * 1e3, 2)   # us per call
        t.fill_(1.0)
# values: random data, many back-to-back calls of changing size, against NCCL (sums of `world` terms: equal to rounding)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
worst = 0.0
same_bits = True
for k, n in enumerate([7, 212, 1 << 20, 3, 4099, 513711, 1, 65536, 16, 182] * 20):
    t = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * (1.0 + k)
    ref = t.clone()
    comm.allreduce(t)
    dist.all_reduce(ref)
    worst = max(worst, float((t - ref).abs().max() / ref.abs().max()))
    if k % 10 == 9:
        allr = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(allr, t)
        same_bits &= all(torch.equal(allr[0], b) for b in allr)
if rank == 0:
    print(json.dumps({"world": dist.get_world_size(), "us_per_call": out, "max_rel_diff_vs_nccl": worst, "same_bits_on_all_ranks": same_bits}))
dist.barrier(); dist.destroy_process_group()
