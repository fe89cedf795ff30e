"""pytest plugin for oracle/run_reference_tests.sh: lets the REFERENCE's own test files (read in place from
/root/reference/tests, build container only) import this package under the name ``astrophot``.

* ``astrophot`` and ``astrophot.<sub>`` resolve to ``astrophot_b200`` and its submodules;
* astropy / matplotlib / pyro / h5py, which the reference's test helpers import and this image lacks, are stubbed
  (as in make_golden.py);
* there is no GPU in the build container, so the device plan is replaced by the oracle-backed stand-in of
  tests/test_lm_host_logic.py -- this checks the HOST side of the drop-in (API surface, control flow, containers),
  not the kernels (tests/ -m gpu does that).
Test infrastructure only."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from make_golden import _Stub  # noqa: E402

sys.meta_path.insert(0, _Stub())

import astrophot_b200  # noqa: E402

sys.modules["astrophot"] = astrophot_b200
for name, mod in list(sys.modules.items()):
    if name.startswith("astrophot_b200."):
        sys.modules["astrophot." + name[len("astrophot_b200."):]] = mod

# our tests' conftest under its own name (test_lm_host_logic imports it), without shadowing the reference's tests
spec = importlib.util.spec_from_file_location("conftest", os.path.join(ROOT, "tests", "conftest.py"))
conftest = importlib.util.module_from_spec(spec)
sys.modules["conftest"] = conftest
spec.loader.exec_module(conftest)

from astrophot_b200 import cabi, fit  # noqa: E402
from test_lm_host_logic import OraclePlan, _solve  # noqa: E402

cabi.Plan = OraclePlan
cabi.lm_solve = _solve
astrophot_b200.AP_config.ap_device = "cpu"
_lm_init = fit.LM.__init__


def _unfused(self, *a, **k):
    k.setdefault("fused_trial", False)      # the stand-in answers the trial's pieces, not the fused device call
    _lm_init(self, *a, **k)


fit.LM.__init__ = _unfused
