"""Trace of the block-sparse PCG on the crowded field: damping, iterations and time of every solve of two short fits."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["bench.py"]
import numpy as np, torch
import bench
import astrophot_b200 as ap
wl = os.environ.get("WL", "c3")
t = bench.build_workload(ap, wl, 1, None)().data.cpu().numpy()
model = bench.build_workload(ap, wl, 1, [bench.make_data(t, 10)])
x0 = bench.start_state(model.parameters.vector_representation().numpy(), scale=bench.start_scale(wl))
lm = ap.fit.LM(model, initial_state=x0, max_iter=6, relative_tolerance=0.0)
orig = lm._solve
trace = []
def traced(L, rhs, loose=False, x0=None):
    n0 = len(lm.pcg_iterations)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h = orig(L, rhs, loose=loose, x0=x0)
    torch.cuda.synchronize()
    its = lm.pcg_iterations[n0:] 
    trace.append({"iter": lm.iteration, "L": L, "loose": bool(loose), "warm": x0 is not None, "its": its, "ms": round(1e3 * (time.perf_counter() - t0), 3)})
    return h
lm._solve = traced
lm.fit()
print(json.dumps({"loss": lm.loss_history, "L_history": lm.L_history}))
for r in trace: print(json.dumps(r))
try:
    import ctypes as C
    from astrophot_b200 import cabi
    buf = (C.c_ulonglong * 16)()
    if cabi.lib().apb_debug_pcg_clk(buf) == 0:
        its = sum(r["its"][0] for r in trace)
        names = ["phase1", "barrier1", "alpha+split", "phase2", "barrier2", "totals"]
        for c in range(2):
            print("cta", "first" if c == 0 else "last", {n: round(buf[8 * c + k] / its / 1e3, 2) for k, n in enumerate(names)}, "us/iteration")
except AttributeError:
    pass
