"""apb_chol_factor / apb_chol_solve against the library calls they replaced (torch.linalg.cholesky_ex -> cuSOLVER potrf,
torch.cholesky_solve -> potrs): microseconds per call, CUDA events, stream-ordered back to back."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from astrophot_b200.cabi import chol_factor, chol_solve


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    out = []
    for P in (200, 500, 1000, 2000, 4000):
        g = torch.Generator(device="cuda").manual_seed(P)
        J = torch.randn(2 * P, P, dtype=torch.float64, device="cuda", generator=g)
        H = J.T @ J
        L = 1.0
        A = H / (1.0 + L)
        d = torch.diagonal(H)
        A.diagonal().copy_(d + L * (1.0 + d))
        rhs = torch.randn(P, dtype=torch.float64, device="cuda", generator=g)
        work, info = chol_factor(H, L)
        chol = torch.linalg.cholesky_ex(A)[0]
        rec = {"P": P,
               "apb_chol_factor_us": timed(lambda: chol_factor(H, L, work=work, info=info)),
               "apb_chol_solve_us": timed(lambda: chol_solve(work, rhs)),
               "torch_build_plus_cholesky_ex_us": timed(lambda: torch.linalg.cholesky_ex((H / (1.0 + L)).diagonal_scatter(d + L * (1.0 + d)))),
               "torch_cholesky_solve_us": timed(lambda: torch.cholesky_solve(rhs.reshape(-1, 1), chol))}
        x = chol_solve(work, rhs)
        rec["rel_residual"] = float((A @ x - rhs).abs().max() / rhs.abs().max())
        out.append(rec)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
