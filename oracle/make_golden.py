"""Generate tests/golden/*.npz by running the REFERENCE itself (CPU torch).

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/make_golden.py [scene ...]``.  The reference's optional
dependencies that the hot path never touches (astropy, matplotlib, pyro, h5py)
are absent here and are stubbed at import.

Per scene the fixture holds the reference's model image(s), its forward-AD
Jacobian (a seeded pixel subset + the full J^T J as a whole-image checksum) in
representation and natural units, and, for LM scenes, the noisy data/variance
and the complete LM history (chi^2, lambda, state per iteration).
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = {"astropy", "matplotlib", "pyro", "h5py"}

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = MagicMock()
        m.__path__, m.__spec__, m.__loader__, m.__name__ = [], spec, self, spec.name
        return m

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _Stub())
    sys.path.insert(0, "/root/reference")
    import astrophot as ap

    ap.AP_config.set_logging_output(stdout=False, filename=None)
    return ap


def _datas(img):
    if hasattr(img, "image_list"):
        return [i.data.detach().cpu().numpy() for i in img.image_list]
    return [img.data.detach().cpu().numpy()]


def main(names):
    import torch
    import scenes

    ap = import_reference()
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in names:
        model, _ = scenes.build(ap, name)
        fix = {}
        x_val = model.parameters.vector_values().detach().cpu().numpy()
        x_rep = model.parameters.vector_representation().detach().cpu().numpy()
        fix["x_val"], fix["x_rep"] = x_val, x_rep
        imgs = _datas(model())
        for i, d in enumerate(imgs):
            fix[f"img{i}"] = d
        rng = np.random.default_rng(99)
        for tag, as_rep in (("rep", True), ("nat", False)):
            J = model.jacobian(as_representation=as_rep)
            Js = _datas(J)
            Jflat = np.concatenate([j.reshape(-1, j.shape[-1]) for j in Js])
            fix[f"jtj_{tag}"] = Jflat.T @ Jflat
            if tag == "rep":
                idx = np.sort(rng.choice(Jflat.shape[0], size=min(3000, Jflat.shape[0]), replace=False))
                fix["jac_idx"] = idx
            fix[f"jac_{tag}"] = Jflat[fix["jac_idx"]]
        if name in scenes.ALL_LM_SCENES:
            seed = scenes.ALL_LM_SCENES[name]
            # embed the truth in full-size target frames (group windows may be smaller)
            tars = model.target.image_list if hasattr(model.target, "image_list") else [model.target]
            wins = model.window.window_list if hasattr(model.window, "window_list") else [model.window]
            full = []
            for t, w, d in zip(tars, wins, imgs):
                f = np.zeros(tuple(t.data.shape))
                f[t.window.get_self_indices(w)] = d
                full.append(f)
            data = scenes.make_data(full, seed, scale=getattr(scenes, "NOISE_SCALE", {}).get(name, 1.0))
            for i, d in data.items():
                fix[f"data{i}"], fix[f"var{i}"] = d["data"], d["variance"]
            m2, _ = scenes.build(ap, name, data=data)
            x0 = scenes.perturb(m2.parameters.vector_representation().detach().cpu().numpy(), seed)
            fix["x0"] = x0
            res = ap.fit.LM(m2, initial_state=x0, max_iter=8, relative_tolerance=0.0, verbose=0)
            # first normal equations (lm.py:256-260) for a direct check
            Y0 = res.forward(parameters=res.current_state).flatten("data")
            J = res.jacobian(parameters=res.current_state).flatten("data")
            fix["hess0"] = res._hess(J, res.W).detach().cpu().numpy()
            fix["grad0"] = res._grad(J, res.W, Y0, res.Y).detach().cpu().numpy()
            fix["ndf"] = np.array(res.ndf, dtype=np.float64)
            res.fit()
            fix["loss_history"] = np.array(res.loss_history)
            fix["L_history"] = np.array(res.L_history)
            fix["lambda_history"] = np.array(res.lambda_history)
            fix["message"] = np.array(res.message)
            # uncertainties in natural units at the fitted state (lm.py:408-425,495-539)
            fix["cov"] = res.covariance_matrix.detach().cpu().numpy()
            res.update_uncertainty()
            fix["uncertainty"] = m2.parameters.vector_uncertainty().detach().cpu().numpy()
            print(name, "LM:", res.message, res.loss_history)
            if name in getattr(scenes, "LM_KWARGS_SCENES", ()):
                # the same fit with non-default LM knobs (geodesic acceleration on, other damping schedule)
                m5, _ = scenes.build(ap, name, data=data)
                r5 = ap.fit.LM(m5, initial_state=x0, max_iter=6, relative_tolerance=0.0, verbose=0,
                               **scenes.LM_KWARGS).fit()
                fix["kw_loss_history"] = np.array(r5.loss_history)
                fix["kw_L_history"] = np.array(r5.L_history)
                fix["kw_lambda_history"] = np.array(r5.lambda_history)
                print(name, "LM kwargs:", r5.message, r5.loss_history)
            if name in getattr(scenes, "ITER_SCENES", ()):
                # fit/iterative.py Iter: 3 sweeps, every sub-fit 4 LM iterations (fixed counts on both sides)
                m3, _ = scenes.build(ap, name, data=data)
                it = ap.fit.Iter(m3, initial_state=x0, max_iter=3,
                                 method_kwargs={"max_iter": 4, "relative_tolerance": 0.0}).fit()
                fix["iter_loss_history"] = np.array(it.loss_history)
                fix["iter_lambda_history"] = np.array(it.lambda_history)
                print(name, "Iter:", it.message, it.loss_history)
                # fit/iterative.py Iter_LM: sequential chunks of 8 parameters, 2 sweeps, 3 LM iterations per chunk;
                # the chunk fits start from the model's parameters
                m4, _ = scenes.build(ap, name, data=data)
                m4.parameters.vector_set_representation(torch.as_tensor(x0, dtype=ap.AP_config.ap_dtype))
                il = ap.fit.Iter_LM(m4, initial_state=x0, chunks=8, method="sequential", max_iter=2,
                                    LM_kwargs={"max_iter": 3, "relative_tolerance": 0.0}).fit()
                fix["iterlm_loss_history"] = np.array(il.loss_history)
                fix["iterlm_lambda_history"] = np.array(il.lambda_history)
                print(name, "Iter_LM:", il.message, il.loss_history)
        path = os.path.join(out_dir, f"{name}.npz")
        np.savez_compressed(path, **fix)
        print(f"wrote {path}: P={len(x_val)} sum={[float(d.sum()) for d in imgs]} "
              f"{os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    import scenes as _s

    main(sys.argv[1:] or _s.SAMPLE_SCENES)
