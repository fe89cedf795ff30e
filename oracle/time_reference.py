"""Time the REFERENCE itself (CPU torch, all cores) on BASELINE config[1] in the build container.
Informational only: /root/reference does not exist on the GPU box, so bench.py's reference arm is
the oracle port.  usage: python oracle/time_reference.py [n_iter]"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import import_reference  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.argv = ["bench.py"]
import bench  # noqa: E402

ap = import_reference()
import astrophot_b200.utils as _u  # noqa: E402  (synthetic-PSF helpers: the reference keeps its own under utils.initialize)
ap.utils.moffat_psf, ap.utils.gaussian_psf = _u.moffat_psf, _u.gaussian_psf
torch.set_num_threads(os.cpu_count())
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
truth_model = bench.build_joint(ap, 1, None)
t0 = time.perf_counter()
truth = truth_model().data.detach().cpu().numpy()
t_sample = time.perf_counter() - t0
datas = [bench.make_data(truth, 10)]
model = bench.build_joint(ap, 1, datas)
x0 = bench.start_state(model.parameters.vector_representation().detach().cpu().numpy(), scale=bench.start_scale("c2"))
t0 = time.perf_counter()
res = ap.fit.LM(model, initial_state=x0, max_iter=n_iter, relative_tolerance=0.0, verbose=0).fit()
dt = time.perf_counter() - t0
its = max(1, len(res.loss_history) - 1)
print(json.dumps({"workload": "c2 (1024^2 Sersic, 51x51 Moffat PSF)", "cores": os.cpu_count(), "sample_s": t_sample,
                  "lm_iterations": its, "s_per_lm_iteration": dt / its, "loss_history": res.loss_history}))
# the reference's LM history at FULL size: fixture for tests/test_cuda_fullsize.py (data are regenerated there from the
# same seeds; the truth images of the two packages agree to 1e-15)
np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "c2_fullsize_lm.npz"), x0=x0,
                    loss_history=np.array(res.loss_history), L_history=np.array(res.L_history),
                    lambda_history=np.array(res.lambda_history), truth_sum=np.array(truth.sum()),
                    truth_probe=truth[::97, ::89].copy())
