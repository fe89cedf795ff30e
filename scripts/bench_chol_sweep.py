"""apb_chol_factor under its tuning knobs (APB_CHOL_RELAXED: barrier poll without acquire; APB_CHOL_GRID: CTAs), one
process: microseconds per factorisation and the residual of a solve with each factor."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astrophot_b200.cabi import chol_factor, chol_solve
from bench_chol import timed

mats = {}
for P in (200, 500, 1000, 2000):
    g = torch.Generator(device="cuda").manual_seed(P)
    J = torch.randn(2 * P, P, dtype=torch.float64, device="cuda", generator=g)
    H = J.T @ J
    A = H / 2.0
    d = torch.diagonal(H)
    A.diagonal().copy_(d + 1.0 * (1.0 + d))
    mats[P] = (H, A, torch.randn(P, dtype=torch.float64, device="cuda", generator=g))
for relaxed in (0, 1):
    for grid in (148, 96, 48, 24):
        os.environ["APB_CHOL_RELAXED"], os.environ["APB_CHOL_GRID"] = str(relaxed), str(grid)
        rec = {"relaxed": relaxed, "grid": grid}
        for P, (H, A, rhs) in mats.items():
            work, info = chol_factor(H, 1.0)
            x = chol_solve(work, rhs)
            res = float((A @ x - rhs).abs().max() / rhs.abs().max())
            assert int(info.item()) == 0 and res < 1e-12, (relaxed, grid, P, res)
            rec[f"P{P}_us"] = round(timed(lambda: chol_factor(H, 1.0, work=work, info=info), n=10))
        print(json.dumps(rec), flush=True)
