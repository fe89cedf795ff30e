"""Run the REFERENCE itself and the numpy oracle on one of bench.py's workloads (same seeded inputs, 2 LM iterations)
and print how far apart their LM histories are.  Build container only (needs /root/reference); this is how the
oracle was checked beyond the small golden scenes:

    python oracle/compare_reference.py c2 | c2b | c3t | c5t | c4band | joint2

Recorded results (chi^2 per iteration, relative difference oracle vs reference; torch 2.11, 8 cores):
    c2      1024^2 Sersic, 51x51 PSF, Jacobian in 2x2 chunks          1.8e-15
    c2b     same with the PSF as a fitted moffat psf model            9.3e-15
    c4band  2048^2 Sersic, 25x25 Gaussian PSF, Jacobian in 3x3 chunks  6e-16
    c3t     512^2 crowded field, 15 Sersic + 78 points + sky, P=340    3e-16
    c5t     512^2 mosaic, Sersic + spline galaxies + sky               2e-16
    joint2  2 bands of c2, shared shape parameters, 8 chunk pieces     7e-16
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
from make_golden import import_reference  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
sys.argv = ["bench.py"]
import bench  # noqa: E402

ref = import_reference()
import astrophot_b200 as ours  # noqa: E402
import astrophot_b200.utils as _u  # noqa: E402
import astrophot_oracle as orc  # noqa: E402
from astrophot_b200.lowering import lower  # noqa: E402

ref.utils.moffat_psf, ref.utils.gaussian_psf = _u.moffat_psf, _u.gaussian_psf
ours.AP_config.ap_device = "cpu"
torch.set_num_threads(os.cpu_count())


def build(ap, datas):
    if wl == "c4band":
        return bench.build_c4_band(ap, datas, 0)
    if wl == "joint2":
        return bench.build_joint(ap, 2, datas)
    return bench.build_workload(ap, wl, 1, datas)


if wl == "joint2":
    tm = bench.build_joint(ref, 1, None)
    datas = []
    for b in range(2):
        tm["Ie"].value = bench.band_truth(b)["Ie"]
        datas.append(bench.make_data(tm().data.detach().cpu().numpy(), 10 + b))
else:
    datas = [bench.make_data(build(ref, None)().data.detach().cpu().numpy(), 10)]
model = build(ref, datas)
scale = 0.05 if wl in ("c2", "c2b", "c4band", "joint2") else 0.02
x0 = bench.start_state(model.parameters.vector_representation().detach().cpu().numpy(), scale=scale)
t0 = time.perf_counter()
res = ref.fit.LM(model, initial_state=x0, max_iter=2, relative_tolerance=0.0, verbose=0).fit()
print(wl, "reference", round(time.perf_counter() - t0, 1), "s", res.loss_history, flush=True)
scene, _ = lower(build(ours, datas), for_fit=True)
orc.set_threads(4)
t0 = time.perf_counter()
r = orc.lm_fit(scene, x0, max_iter=2, relative_tolerance=0.0, conv="fft")
print(wl, "oracle   ", round(time.perf_counter() - t0, 1), "s", r["loss_history"], flush=True)
n = min(len(r["loss_history"]), len(res.loss_history))
print(wl, "chi^2 rel diff", np.abs(np.array(r["loss_history"][:n]) / np.array(res.loss_history[:n]) - 1).max(),
      "state diff", np.abs(np.array(r["lambda_history"][:n]) - np.array(res.lambda_history[:n])).max())
