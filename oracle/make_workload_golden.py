"""LM histories of the REFERENCE itself on bench.py workloads at sizes the reference can run (build container only:
needs /root/reference).  Fixtures for tests/test_cuda_fullsize.py:

    python oracle/make_workload_golden.py c5t      512^2 scale model of BASELINE config[4]: Sersic + spline galaxies + sky
    python oracle/make_workload_golden.py joint8   BASELINE config[3] at 160^2 per band: 8 bands, shared centre / q / PA / n / Re,
                                                  per-band Ie and Gaussian PSF (Target_Image_List)
    python oracle/make_workload_golden.py c3t | c4band   (the fixtures of round 1, same recipe)

Data are regenerated in the tests from the same seeds (bench.make_data); the truth images of the two packages agree to
1e-15, the fixture keeps the sum and a probe of the reference's truth to check that."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
from make_golden import import_reference  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

wl = sys.argv[1]
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sys.argv = ["bench.py"]
import bench  # noqa: E402

JOINT8_SIZE = 160

ref = import_reference()
import astrophot_b200.utils as _u  # noqa: E402
ref.utils.moffat_psf, ref.utils.gaussian_psf = _u.moffat_psf, _u.gaussian_psf
torch.set_num_threads(os.cpu_count())

if wl == "joint8":
    truth = []
    for b in range(bench.C4_BANDS):
        full = bench.build_c4(ref, None, size=JOINT8_SIZE)       # (a band's truth = its sub-model sampled alone)
        truth.append(list(full.models.values())[b]().data.detach().cpu().numpy())
    datas = [bench.make_data(t, 10 + b) for b, t in enumerate(truth)]
    model = bench.build_c4(ref, datas, size=JOINT8_SIZE)
    scale = bench.start_scale("c4")
elif wl == "c4band":
    truth = [bench.build_c4_band(ref, None, 0)().data.detach().cpu().numpy()]
    datas = [bench.make_data(truth[0], 10)]
    model = bench.build_c4_band(ref, datas, 0)
    scale = bench.start_scale("c4")
else:
    truth = [bench.build_workload(ref, wl, 1, None)().data.detach().cpu().numpy()]
    datas = [bench.make_data(truth[0], 10)]
    model = bench.build_workload(ref, wl, 1, datas)
    scale = bench.start_scale(wl)
x0 = bench.start_state(model.parameters.vector_representation().detach().cpu().numpy(), scale=scale)
res = ref.fit.LM(model, initial_state=x0, max_iter=n_iter, relative_tolerance=0.0, verbose=0).fit()
print(wl, res.loss_history, res.L_history, res.message)
out = os.path.join(os.path.dirname(HERE), "tests", "golden", f"{wl}_lm.npz")
np.savez_compressed(out, x0=x0, loss=np.array(res.loss_history), L=np.array(res.L_history),
                    lam=np.array(res.lambda_history), truth_sum=np.array([t.sum() for t in truth]),
                    truth_probe=np.concatenate([t[::37, ::41].reshape(-1) for t in truth]))
print("wrote", out)
