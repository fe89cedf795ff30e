#!/usr/bin/env python
"""Benchmark of the AstroPhot forward-model-and-fit hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

Metric (BASELINE.json): LM iterations/s (model + PSF + J^T J), fp64.  A *step* is one Levenberg-Marquardt iteration
(fit/lm.py:450-467): one fused sample+Jacobian+normal-equation build and k lambda-trials, each with a damped solve, a
geodesic pass and a chi^2 pass.  Steps follow fits of 10 LM iterations from a perturbed start (a fit also
ends when chi^2 has reached its noise floor or LM cannot improve it), one after the other, so K steps are K real
iterations; the CPU arm walks the same fits.  The K-step block is repeated until 2 s have been timed and the median block
is reported.

Headline workload (default, every N): `c3` = BASELINE config[2], the crowded field the north star's 1-GPU target is
quoted on -- 1000 PSF-convolved Sersic + 5000 point sources + sky on 4096x4096, P = 22001.  At N > 1 the image is cut
into N tiles, one per GPU (strong scaling; `value` = LM iterations/s of the one fit); the ranks exchange the block-sparse
J^T W J, J^T W r and chi^2 (NCCL all-reduce).  The other configurations are measured in the same run with shorter timed
regions and attached to the line as `other_workloads`: `c4` = config[3], the 8-band joint fit on 2048^2 per band
(8/N bands per GPU, strong scaling), `c2` = config[1] (one 1024^2 band per GPU, weak scaling) and, at N = 1, `c5s`, the
2048^2 scale model of config[4].  `--workload X` measures X alone.

`--impl reference` times the CPU oracle port of the reference algorithm (oracle/astrophot_oracle.py, numpy + scipy FFT
convolution like the reference's default psf_convolve_mode) on the host cores with the same iteration mix; the crowded
field on its 512^2 scale model c3t (the dense Jacobian of c3 itself would need 2.9 TB), extrapolated linearly in the
source count (factor 64, stated in `config`).  /root/reference is not needed at run time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "lm_iters_per_sec"
UNIT = "LM iterations/s (x bands)"
SIZE = 1024
PSF_W = 51


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def band_truth(b):
    return {"center": [SIZE / 2 + 0.3, SIZE / 2 - 0.4], "q": 0.6, "PA": 1.0, "n": 2.5, "Re": 60.0, "Ie": 1.0 + 0.1 * b}


def build_joint(ap, n_bands, datas, size=SIZE, aux_psf=False):
    """n_bands PSF-convolved Sersic models sharing shape parameters
    (docs/source/tutorials/JointModels.ipynb recipe).  datas[b] = dict(data, variance) or None.
    aux_psf: the PSF is a `moffat psf model` (n = 2.5, Rd = 3.0 on a 51x51 PSF_Image grid) fitted together with
    the galaxy (SURVEY.md §8d C2 variant B, P = 9) instead of a fixed PSF_Image."""
    psf_np = ap.utils.moffat_psf(2.5, 3.0, PSF_W, 1.0)
    tars, models = [], []
    for b in range(n_bands):
        d = datas[b] if datas is not None else None
        kw = {} if d is None else {"variance": d["variance"]}
        if not aux_psf:
            kw["psf"] = ap.image.PSF_Image(data=psf_np, pixelscale=1.0)
        tars.append(ap.image.Target_Image(data=np.zeros((size, size)) if d is None else d["data"], pixelscale=1.0,
                                          zeropoint=22.5, **kw))
    for b in range(n_bands):
        pars = band_truth(b)
        pars["center"] = [size / 2 + 0.3, size / 2 - 0.4]
        kw = {}
        if aux_psf:
            ptar = ap.image.PSF_Image(data=np.zeros((PSF_W, PSF_W)), pixelscale=1.0)
            kw["psf"] = ap.models.AstroPhot_Model(name=f"psf{b}", model_type="moffat psf model", target=ptar,
                                                  parameters={"n": 2.5, "Rd": 3.0})
        m = ap.models.AstroPhot_Model(name=f"band{b}", model_type="sersic galaxy model", target=tars[b],
                                      psf_mode="full", parameters=pars, **kw)
        if b > 0:
            for p in ("center", "q", "PA", "n", "Re"):
                m[p].value = models[0][p]
        models.append(m)
    if n_bands == 1:
        return models[0]
    return ap.models.AstroPhot_Model(name="joint", model_type="group model", models=models,
                                     target=ap.image.Target_Image_List(tars), psf_mode="full")


C4_BANDS, C4_SIZE, C4_PSF = 8, 2048, 25


def build_c4(ap, datas, size=C4_SIZE):
    """BASELINE config[3]: 8 bands, shared centre / q / PA / n / Re, per-band Ie = 0.3 + 0.1 b,
    per-band Gaussian PSF 25x25 with sigma = 1.2 + 0.1 b px, psf_mode full, Target_Image_List."""
    tars, models = [], []
    for b in range(C4_BANDS):
        d = datas[b] if datas is not None else None
        kw = {} if d is None else {"variance": d["variance"]}
        psf_np = ap.utils.gaussian_psf(1.2 + 0.1 * b, C4_PSF, 1.0)
        tars.append(ap.image.Target_Image(data=np.zeros((size, size)) if d is None else d["data"], pixelscale=1.0,
                                          zeropoint=22.5, psf=ap.image.PSF_Image(data=psf_np, pixelscale=1.0), **kw))
    for b in range(C4_BANDS):
        pars = {"center": [size / 2 + 0.3, size / 2 - 0.4], "q": 0.6, "PA": 1.0, "n": 2.5, "Re": 60.0 * size / 1024,
                "Ie": 0.3 + 0.1 * b}
        m = ap.models.AstroPhot_Model(name=f"band{b}", model_type="sersic galaxy model", target=tars[b],
                                      psf_mode="full", parameters=pars)
        if b > 0:
            for p in ("center", "q", "PA", "n", "Re"):
                m[p].value = models[0][p]
        models.append(m)
    return ap.models.AstroPhot_Model(name="joint8", model_type="group model", models=models,
                                     target=ap.image.Target_Image_List(tars), psf_mode="full")


TILES = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}         # image tiles of the crowded field at N GPUs
if os.environ.get("APB_TILES"):      # experiment: e.g. APB_TILES=4x8 deals 32 tiles to the ranks round robin
    _ty, _tx = (int(v) for v in os.environ["APB_TILES"].split("x"))
    TILES = {n: (_ty, _tx) for n in TILES}
C3 = {"c3": (4096, 1000, 5000), "c3s": (1024, 62, 312), "c3t": (512, 15, 78)}     # size, Sersic sources, point sources


def build_crowded(ap, workload, data=None):
    """BASELINE config[2]: crowded field, PSF-convolved Sersic sources (128^2 windows) + point sources
    (53^2 windows) + flat sky on one image with a 51x51 Moffat PSF.  `c3s` is the 1024^2 scale model
    with the same source densities (the largest the reference's dense Jacobian fits in host memory)."""
    size, n_gal, n_pt = C3[workload]
    rng = np.random.default_rng(3)
    psf_np = ap.utils.moffat_psf(2.5, 2.0, PSF_W, 1.0)
    kw = {} if data is None else {"variance": data["variance"]}
    tar = ap.image.Target_Image(data=np.zeros((size, size)) if data is None else data["data"], pixelscale=1.0,
                                zeropoint=22.5, psf=ap.image.PSF_Image(data=psf_np, pixelscale=1.0), **kw)
    M = ap.models.AstroPhot_Model
    models = []

    def box(c, lo, hi):
        return [max(0, int(c) - lo), min(size, int(c) + hi)]

    for k in range(n_gal):
        cx, cy = rng.uniform(40, size - 40, size=2)
        models.append(M(name=f"g{k}", model_type="sersic galaxy model", target=tar, psf_mode="full",
                        window=[box(cx, 64, 64), box(cy, 64, 64)],
                        parameters={"center": [cx, cy], "q": rng.uniform(0.4, 0.9), "PA": rng.uniform(0, np.pi),
                                    "n": rng.uniform(1, 4), "Re": rng.uniform(3, 12), "Ie": rng.uniform(0, 1)}))
    for k in range(n_pt):
        cx, cy = rng.uniform(40, size - 40, size=2)
        models.append(M(name=f"p{k}", model_type="point model", target=tar, window=[box(cx, 26, 27), box(cy, 26, 27)],
                        parameters={"center": [cx, cy], "flux": rng.uniform(1, 2)}))
    sky = M(name="sky", model_type="flat sky model", target=tar, parameters={"F": {"value": -2.0, "uncertainty": 0.01}})
    sky.initialize()
    models.append(sky)
    return M(name="crowd", model_type="group model", models=models, target=tar, psf_mode="full")


C5 = {"c5": (16384, 10000), "c5s": (2048, 156), "c5t": (512, 10)}     # size, galaxies
C5_PROF = [0.0, 1.0, 2.0, 3.0, 4.5, 6.5, 9.0, 12.0, 16.0, 22.0, 30.0, 42.0]


def build_mosaic(ap, workload, data=None):
    """BASELINE config[4]: low-surface-brightness mosaic, galaxies in 96^2 windows (70 % Sersic, 30 % spline
    galaxy models with 12 radii) plus a flat sky, no PSF.  `c5s` / `c5t` are scale models with the same density."""
    size, n_gal = C5[workload]
    rng = np.random.default_rng(5)
    kw = {} if data is None else {"variance": data["variance"]}
    tar = ap.image.Target_Image(data=np.zeros((size, size)) if data is None else data["data"], pixelscale=1.0,
                                zeropoint=22.5, **kw)
    M = ap.models.AstroPhot_Model
    models = []
    prof = np.array(C5_PROF)
    for k in range(n_gal):
        cx, cy = rng.uniform(50, size - 50, size=2)
        win = [[int(cx) - 48, int(cx) + 48], [int(cy) - 48, int(cy) + 48]]
        q, pa, n, re, ie = rng.uniform(0.4, 0.9), rng.uniform(0, np.pi), rng.uniform(1, 4), rng.uniform(3, 12), rng.uniform(0, 1)
        if rng.uniform() < 0.7:
            models.append(M(name=f"g{k}", model_type="sersic galaxy model", target=tar, window=win,
                            parameters={"center": [cx, cy], "q": q, "PA": pa, "n": n, "Re": re, "Ie": ie}))
        else:
            bn = 2 * n - 1 / 3
            val = ie - bn * ((np.maximum(prof, 0.05) / re) ** (1 / n) - 1) / np.log(10)
            models.append(M(name=f"s{k}", model_type="spline galaxy model", target=tar, window=win,
                            parameters={"center": [cx, cy], "q": q, "PA": pa,
                                        "I(R)": {"value": [float(v) for v in val], "prof": [float(r) for r in prof]}}))
    sky = M(name="sky", model_type="flat sky model", target=tar, parameters={"F": {"value": -2.0, "uncertainty": 0.01}})
    sky.initialize()
    models.append(sky)
    return M(name="mosaic", model_type="group model", models=models, target=tar)


def build_workload(ap, workload, n_bands, datas):
    if workload in C5:
        return build_mosaic(ap, workload, None if datas is None else datas[0])
    if workload == "c2":
        return build_joint(ap, n_bands, datas)
    if workload == "c2b":
        return build_joint(ap, 1, datas, aux_psf=True)
    if workload == "c4":
        return build_c4(ap, datas) if n_bands > 1 else build_c4_band(ap, datas)
    return build_crowded(ap, workload, None if datas is None else datas[0])


def build_c4_band(ap, datas, b=0):
    """One band of c4 on its own (truth images are sampled band by band)."""
    d = datas[0] if datas is not None else None
    kw = {} if d is None else {"variance": d["variance"]}
    psf_np = ap.utils.gaussian_psf(1.2 + 0.1 * b, C4_PSF, 1.0)
    tar = ap.image.Target_Image(data=np.zeros((C4_SIZE, C4_SIZE)) if d is None else d["data"], pixelscale=1.0,
                                zeropoint=22.5, psf=ap.image.PSF_Image(data=psf_np, pixelscale=1.0), **kw)
    pars = {"center": [C4_SIZE / 2 + 0.3, C4_SIZE / 2 - 0.4], "q": 0.6, "PA": 1.0, "n": 2.5, "Re": 60.0 * C4_SIZE / 1024,
            "Ie": 0.3 + 0.1 * b}
    return ap.models.AstroPhot_Model(name=f"band{b}", model_type="sersic galaxy model", target=tar, psf_mode="full",
                                     parameters=pars)


def workload_text(workload, n_bands, world=1):
    if workload == "c4":
        return (f"c4: {C4_BANDS}-band joint fit (shared centre/q/PA/n/Re, per-band Ie), PSF-convolved Sersic on "
                f"{C4_SIZE}x{C4_SIZE} per band, {C4_PSF}x{C4_PSF} Gaussian PSF per band, threshold sub-pixel integration, "
                f"LM fp64, {C4_BANDS // world} band(s) per GPU")
    if workload in C5:
        size, n_gal = C5[workload]
        return (f"{workload}: LSB mosaic, {n_gal} galaxies in 96^2 windows (70 % Sersic, 30 % spline with 12 radii) + flat sky "
                f"on {size}x{size}, no PSF, threshold sub-pixel integration, LM fp64"
                + (f", image cut into {TILES[world][0]}x{TILES[world][1]} tiles, one per GPU" if world > 1 else ""))
    if workload == "c2b":
        return (f"c2b: PSF-convolved Sersic on {SIZE}x{SIZE}, PSF = moffat psf model on a {PSF_W}x{PSF_W} grid fitted as "
                "auxiliary parameters (P = 9), threshold sub-pixel integration, LM fp64")
    if workload == "c2":
        return (f"c2 x {n_bands} band(s): PSF-convolved Sersic, {SIZE}x{SIZE} per band, {PSF_W}x{PSF_W} Moffat PSF, "
                "threshold sub-pixel integration, LM fp64, joint fit sharded 1 band/GPU")
    size, n_gal, n_pt = C3[workload]
    return (f"{workload}: crowded field, {n_gal} PSF-convolved Sersic (128^2 windows) + {n_pt} point sources (53^2 windows) "
            f"+ flat sky on {size}x{size}, {PSF_W}x{PSF_W} Moffat PSF, threshold sub-pixel integration, LM fp64"
            + (f", image cut into {TILES[world][0]}x{TILES[world][1]} tiles, one per GPU" if world > 1 else ""))


def make_data(truth, seed):
    rng = np.random.default_rng(seed)
    var = 0.1**2 + truth / 100.0
    return {"data": truth + rng.normal(size=truth.shape) * np.sqrt(var), "variance": var}


def start_state(x_rep, seed=2, scale=0.05):
    rng = np.random.default_rng(1000 + seed)
    return np.asarray(x_rep, dtype=np.float64) + scale * rng.normal(size=len(x_rep))


def start_scale(workload):
    return 0.05 if workload in ("c2", "c2b", "c4") else 0.02


# ---------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10.0:     # nvidia-smi needs ~1 s to print its first row
                time.sleep(0.05)
        except Exception:
            self.proc = None
        return self

    def mark(self):
        """Index of the next sample: samples[mark0:mark1] were taken inside a region."""
        return len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, lo=0, hi=None):
        sm, mx, reasons = [], [], set()
        rows = self.rows[lo:hi]
        if not rows:                       # region shorter than one sampling period: nearest samples
            rows = self.rows[max(0, lo - 1):(hi or len(self.rows)) + 1]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# reference arm: CPU oracle port
# ---------------------------------------------------------------------------
FIT_ITERS = 10         # fits of a fixed number of LM iterations from the perturbed start, both arms
FLOOR_TOL = 1e-9       # ... ended early when chi^2 has reached its noise floor (or LM cannot improve it: OptimizeStop)


def converged(loss, L=None):
    """Both arms start the next fit after FIT_ITERS iterations, or as soon as an iteration moved chi^2 by less than
    FLOOR_TOL: past that point an LM iteration only burns lambda-trials on the rounding noise of chi^2, and how many is
    decided by the last bit of a sum (it differs between summation orders of the same all-reduce)."""
    n = len(loss) - 1
    return n >= FIT_ITERS or (n >= 1 and abs(loss[-2] - loss[-1]) / loss[-1] < FLOOR_TOL)


def cpu_lm_iterations(scene, x0, n_skip, n_timed, budget_s=None):
    """Seconds of each of `n_timed` consecutive LM iterations of the oracle port, taken after `n_skip` untimed ones, along
    the SAME trajectory the GPU arm walks: the fit starts from the perturbed state x0, continues iteration after
    iteration, and starts over from x0 whenever it converges (`converged`) or cannot improve chi^2 any more."""
    import astrophot_oracle as orc

    orc.set_threads(os.cpu_count())      # sources / convolution planes on all host cores; BLAS threads for J^T W J
    times, t_start, restarts = [], time.perf_counter(), 0
    while len(times) < n_skip + n_timed:
        res = orc.lm_fit(scene, x0, max_iter=n_skip + n_timed - len(times), relative_tolerance=0.0, conv="fft",
                         stop=converged)
        times += list(res["iter_seconds"])
        if len(times) < n_skip + n_timed:
            restarts += 1
        if not res["iter_seconds"]:
            break
        if budget_s is not None and len(times) > n_skip and time.perf_counter() - t_start > budget_s:
            break
    return times[n_skip:n_skip + n_timed], restarts


def cpu_scene(workload="c2"):
    """Scene tables for the CPU arm (host numpy only, no GPU needed).  The crowded field is timed on its
    scale model c3t: the reference's (and the port's) dense Jacobian of c3 itself would need 2.9 TB."""
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    from astrophot_b200.lowering import lower

    if workload in ("c3", "c3s"):
        workload = "c3t"
    if workload in ("c5", "c5s"):
        workload = "c5t"
    dev = ap.AP_config.ap_device
    ap.AP_config.ap_device = "cpu"
    try:
        model = build_workload(ap, workload, 1, None)
        scene, _ = lower(model)
        xv = model.parameters.vector_values().numpy()
        truth = orc.sample(scene, xv, as_rep=False, conv="fft")
        datas = [make_data(t, 10 + b) for b, t in enumerate(truth)]
        model = build_workload(ap, workload, 1, datas)
        scene, _ = lower(model, for_fit=True)
        x0 = start_state(model.parameters.vector_representation().numpy(), scale=start_scale(workload))
    finally:
        ap.AP_config.ap_device = dev
    return scene, x0


def cpu_scale(workload):
    """(factor, note, scale-model description): the crowded field is timed on its 512^2 scale model c3t (same source
    densities; c3 is 64 x, c3s 4 x c3t in sources and pixels) and the CPU cost per LM iteration is taken as linear in that
    (it is super-linear for the dense J^T W J, so this flatters the CPU); c4 is timed on one of its 8 bands."""
    if workload in ("c3", "c3s"):
        k = 64 if workload == "c3" else 4
        return 1.0 / k, (f" on the scale model c3t (512^2, 15 Sersic + 78 points + sky, P = 340), divided by {k} "
                         f"(linear extrapolation to {workload}: EXTRAPOLATED)"), {"scale_model": "c3t", "extrapolation_factor": k}
    if workload in ("c5", "c5s"):
        k = 1024 if workload == "c5" else 16
        return 1.0 / k, (f" on the scale model c5t (512^2, 10 galaxies + sky), divided by {k} "
                         f"(linear extrapolation to {workload}: EXTRAPOLATED)"), {"scale_model": "c5t", "extrapolation_factor": k}
    if workload == "c4":
        return 1.0 / C4_BANDS, (f" on ONE band of the {C4_BANDS}-band joint fit (2048^2, P = 7), divided by {C4_BANDS} "
                                "(EXTRAPOLATED)"), {"scale_model": "one band of c4", "extrapolation_factor": C4_BANDS}
    return 1.0, "", {"scale_model": None, "extrapolation_factor": 1}


def default_workload(world):
    """The headline workload: BASELINE config[2], the 4096^2 crowded field -- the configuration the north star's 1-GPU
    target is quoted on -- at every N (N > 1: the image cut into N tiles, strong scaling)."""
    return "c3"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    wl = args.workload or default_workload(args.gpus)
    scene, x0 = cpu_scene(wl)
    fac, note, scale_info = cpu_scale(wl)
    # the same iteration mix as the GPU arm: W untimed iterations from the perturbed start, then K timed ones along the
    # same trajectory, fits of FIT_ITERS iterations (`converged`).  The run is bounded by a time budget, so fewer than K
    # iterations may be timed (steps_timed says how many).
    budget = float(os.environ.get("APB_REF_BUDGET_S", "170"))
    times, restarts = cpu_lm_iterations(scene, x0, args.warmup, args.steps, budget_s=budget)
    ms = 1e3 * float(np.mean(times)) / fac
    val = 1e3 / ms
    crowded = wl in C3 or wl in C5
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "steps_timed": len(times), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if wl == "c2" else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": workload_text(wl, args.gpus if wl == "c2" else 1, args.gpus if crowded or wl == "c4" else 1),
                        "fit_restarts": restarts,
                        "note": "CPU oracle port of the reference algorithm (numpy + scipy FFT conv); LM iterations along the fit "
                                "trajectory from the perturbed start, restarted on convergence: the GPU arm's iteration mix" + note},
                       **scale_info),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} consecutive LM iterations (1 normal-equation build + lambda trials each) after "
                                   f"{args.warmup} untimed ones; sources and convolution planes on a thread pool of all cores, BLAS "
                                   "threads for J^T W J (the numpy port is mostly serial: ~1.2x over one thread)" + note},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# algorithmic work of the kernels, summed over the profiled region
# ---------------------------------------------------------------------------
def _fft_len(n):
    from astrophot_b200 import cabi
    return cabi.fft_length(n)


# FP64-pipe instructions per profile evaluation (SPE), Sersic:
#   NOMINAL: SURVEY.md §8d, reference-faithful pow form -- 138 value-only, 230 value + 7 derivatives;
#   EXECUTED: what this implementation issues per SPE (table-driven exp / log, rotation once per cell), counted from the
#   ncu source page of the integration kernels (profiles/r02_summary.md): 40 value-only, 95 with derivatives.
# `frac` of an FP64-bound kernel uses the EXECUTED count (it then reads as FP64-pipe utilisation and can be checked
# against ncu's sm__pipe_fp64_cycles_active); `frac_nominal` uses the SURVEY figure and may exceed 1.
SPE_NOMINAL = {"v": 138.0, "g": 230.0}
SPE_EXECUTED = {"v": 40.0, "g": 95.0}


def algorithmic_work(scene, spe, kern, n_fwd, n_jac, n_geo=0):
    """Algorithmic bytes / FP64 instructions of each kernel over the profiled region, summed over the sources of the
    scene.  n_fwd value-only sampling passes, n_jac value+derivative passes, n_geo geodesic J^T v passes;
    spe = {"first": [value-only, derivative], "queue": [value-only, derivative]} profile evaluations the first pass and
    the refinement queues really did in that region (apb_plan_stats totals)."""
    from astrophot_b200 import scene as sc

    tot = lambda f, j: n_fwd * f + n_jac * j
    fft_rows = fft_cols = fft_inv = conv_flops = blocks = 0.0
    n_fft = n_dir = 0
    for src in scene.sources:
        n_act = sum(1 for sl in src.slot if sl >= 0)
        ow, oh = src.out[2], src.out[3]
        px = ow * oh
        if src.kind not in (sc.KIND_POINT, sc.KIND_FLAT_SKY):
            # (n_act derivative planes + weight + residual) x 8 B + mask per pixel per build; J^T v: planes + v + mask
            blocks += n_jac * px * ((n_act + 2) * 8 + 1) + n_geo * px * ((n_act + 1) * 8 + 1)
        if src.psf < 0 or src.kind in (sc.KIND_POINT, sc.KIND_FLAT_SKY):
            continue
        ps = scene.psfs[src.psf]
        pw = int(ps.data.shape[1]) if ps.data is not None else int(ps.shape[1])
        shifted = src.psf_shift != 0
        spw = pw + (2 if shifted else 0)      # bilinear-shifted stamp keeps its 1-px pad
        b = (pw + 2) // 2                     # psf_border_int = ceil((P+1)/2)
        ew, eh = ow + 2 * b, oh + 2 * b
        # planes per pass: value-only: 1 image plane, 1 PSF plane, 1 product; derivative pass: value + (n_act-2)
        # non-centre planes in, 3 PSF planes (K, dK/dcx, dK/dcy), 1 + n_act products out
        in_j, k_j, j_j = 1 + max(n_act - 2, 0), 3, 1 + n_act
        if spw * spw > 17 * 17:
            n_fft += 1
            nx, ny = _fft_len(ew), _fft_len(eh)
            nxh = nx // 2 + 1
            rows = lambda n_in, n_k: n_in * (eh * ew * 8 + eh * nxh * 16) + n_k * (spw * spw * 8 + spw * nxh * 16)
            cols = lambda n_k, n_j: n_k * (spw * nxh * 16 + nxh * ny * 16) + n_j * (eh * nxh * 16 + nxh * ny * 16 + oh * nxh * 16)
            inv = lambda n_j: n_j * (oh * nxh * 16 + px * 8)
            fft_rows += tot(rows(1, 1), rows(in_j, k_j))
            fft_cols += tot(cols(1, 1), cols(k_j, j_j))
            fft_inv += tot(inv(1), inv(j_j))
        else:
            n_dir += 1
            conv_flops += 2.0 * spw * spw * px * tot(1, j_j)
    # k_integrate and k_integrate_pool are launched back to back; the one whose kind of queue it is not returns at once
    pooled = kern.get("k_integrate_pool", (0, 0.0))[1] > kern.get("k_integrate", (0, 0.0))[1]
    k_int, k_int_g = ("k_integrate_pool", "k_integrate_pool_grad") if pooled else ("k_integrate", "k_integrate_grad")

    def fp64(n_spe, kind, what):
        return {"bound": "fp64", "flops": 2.0 * SPE_EXECUTED[kind] * n_spe, "flops_nominal": 2.0 * SPE_NOMINAL[kind] * n_spe,
                "spe": n_spe, "what": f"{n_spe:.0f} profile evaluations x {SPE_EXECUTED[kind]:.0f} FP64 instr executed "
                                      f"(x2 flop; nominal {SPE_NOMINAL[kind]:.0f}): {what}"}

    work = {
        "k_conv": {"bound": "fp64", "flops": conv_flops, "flops_nominal": conv_flops,
                   "what": f"2*P_s^2 flop per output pixel and plane, {n_dir} direct-convolved sources"},
        "k_fft_rows": {"bound": "hbm", "bytes": float(fft_rows),
                       "what": f"real rows in (8 B/px) + half spectra out (16 B x nxh/row), {n_fft} FFT-convolved sources"},
        "k_fft_cols": {"bound": "hbm", "bytes": float(fft_cols),
                       "what": "per product: read eh x nxh spectrum + nxh x Ny PSF spectrum, write oh x nxh; 16 B each"},
        "k_fft_rows_inv": {"bound": "hbm", "bytes": float(fft_inv), "what": "half spectra in (16 B x nxh/row) + real rows out (8 B/px)"},
        "k_blocks": {"bound": "hbm", "bytes": float(blocks),
                     "what": "J^T W J build: n_act derivative planes + weight + residual (8 B) + mask per pixel of every source window; "
                             "geodesic J^T v: n_act planes + v + mask per pixel (pair blocks not counted)"},
        "k_first": fp64(spe["first"][0], "v", "first-pass evaluations of the value-only passes"),
        "k_first_grad": fp64(spe["first"][1], "g", "first-pass evaluations of the derivative passes"),
        k_int: fp64(spe["queue"][0], "v", "Gauss-Legendre nodes of every refinement-queue entry, all depths, value-only passes"),
        k_int_g: fp64(spe["queue"][1], "g", "Gauss-Legendre nodes of every refinement-queue entry, derivative passes"),
    }
    return {k: v for k, v in work.items() if v.get("flops", 0) or v.get("bytes", 0)}


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def measure(wl, args, ctx, main=True):
    """Every number of one workload at the current world size: device-resident value, per-kernel times + roofline,
    end-to-end value; the dict is (for the main workload) the bench line."""
    import torch
    import torch.distributed as dist
    import astrophot_b200 as ap
    from astrophot_b200 import cabi
    from astrophot_b200.errors import OptimizeStop

    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    crowded = wl in C3 or wl in C5        # one big image: cut into tiles at N > 1
    if wl == "c2b" and world > 1:
        raise SystemExit("c2b is a single-band fit (replicas only)")
    if crowded and world not in TILES:
        raise SystemExit("the crowded field is cut into 1, 2, 4 or 8 tiles")
    if wl == "c4" and C4_BANDS % world:
        raise SystemExit("c4 has 8 bands: --gpus must divide 8")
    # c2: one band per GPU (weak scaling); c4: 8 bands dealt to the GPUs; c3: one image cut into `world` tiles (strong)
    n_bands = world if wl == "c2" else (C4_BANDS if wl == "c4" else 1)   # c2b, c3*, c5*: one image
    scaling = "weak" if wl == "c2" else "strong"
    units_per_step = n_bands if wl == "c2" else 1
    steps, warmup = args.steps, args.warmup
    min_timed_s = 2.0 if main else 0.5

    # truth + noisy data; every rank builds all descriptions, data only for the bands it owns (the crowded field's
    # one image is built whole on every rank and cut by LM(tiles=...))
    datas = []
    if crowded:
        t = build_workload(ap, wl, 1, None)().data.cpu().numpy()
        datas.append(make_data(t, 10))
    else:
        truth_model = build_workload(ap, wl, 1, None) if wl in ("c2", "c2b") else None
        for b in range(n_bands):
            if b % world == rank:
                if wl in ("c2", "c2b"):
                    truth_model["Ie"].value = band_truth(b)["Ie"]
                    t = truth_model().data.cpu().numpy()
                else:
                    t = build_c4_band(ap, None, b)().data.cpu().numpy()
                datas.append(make_data(t, 10 + b))
            else:
                datas.append(None)
        del truth_model
    torch.cuda.empty_cache()
    model = build_workload(ap, wl, n_bands, datas)
    x_true = model.parameters.vector_representation().numpy()
    x0 = start_state(x_true, scale=start_scale(wl))
    lm = ap.fit.LM(model, initial_state=x0, max_iter=10**6, relative_tolerance=0.0, distributed=(world > 1),
                   conv=args.conv, tiles=(TILES[world] if crowded and world > 1 else None))
    plan = lm.plan
    n_pix_local = sum(h * w for h, w in plan.shapes)
    flush = ctx["flush"]

    state = {"iters_in_fit": 0, "restarts": 0}

    def reset():
        lm.current_state = torch.as_tensor(x0, dtype=torch.float64, device=dev)
        lm.L = 1.0
        lm.loss_history = [lm._chi2_record(lm.current_state)]
        lm.L_history, lm.lambda_history = [lm.L], [x0.copy()]
        state["iters_in_fit"] = 0

    def one_iteration():
        """One pass of the `for iteration in range(max_iter)` loop of LM.fit (fit/lm.py:450-467)."""
        try:
            res = lm.step(chi2=lm.loss_history[-1])
        except OptimizeStop:
            state["restarts"] += 1
            reset()
            res = lm.step(chi2=lm.loss_history[-1])
        lm.L = res[2]
        lm.current_state = (lm.current_state + res[0]).detach()
        lm.L_history.append(lm.L)
        lm.loss_history.append(res[1])
        lm.Ldn()
        state["iters_in_fit"] += 1
        if converged(lm.loss_history[-(state["iters_in_fit"] + 1):]):
            state["restarts"] += 1
            reset()     # converged: start the next fit (outside the next step's timing)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ctx["clocks"]
    reset()
    m0 = clocks.mark()
    for _ in range(warmup):
        one_iteration()
    barrier()

    # ---- timed region (device-resident inputs): blocks of K iterations, L2 flushed between iterations; the block is
    #      repeated until min_timed_s have been timed and the MEDIAN block is reported (every rank runs the same count)
    def timed_block():
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for k in range(steps):
            flush.zero_()
            barrier()
            ev[k][0].record()
            one_iteration()
            ev[k][1].record()
        barrier()
        return allmax(float(sum(a.elapsed_time(b) for a, b in ev)))     # max over ranks

    launches0 = cabi.launch_count()
    blocks_ms = [timed_block()]
    while sum(blocks_ms) < 1e3 * min_timed_s and len(blocks_ms) < 400:
        blocks_ms.append(timed_block())
    launches = cabi.launch_count() - launches0
    total_ms = float(np.median(blocks_ms))
    ms_per_step = total_ms / steps
    value = units_per_step * steps / (total_ms * 1e-3)

    # ---- the same K iterations again with every kernel launch bracketed by CUDA events on its stream
    #      (per-kernel durations for the roofline; kept out of `value` because the event records cost time)
    reset()
    for _ in range(warmup):
        one_iteration()
    plans = lm.all_plans     # main plan, chi^2 twin, speculative pair
    spec = lm.speculate
    lm.speculate = False     # per-kernel durations are taken without the concurrent guess trials
    for pl in plans:
        pl.profile(True)
        pl.profile_read(reset=True)
    trials0, fwd0, jac0 = lm.n_trials, lm.n_forward, lm.n_jacobian
    st0 = [pl.stats() for pl in plans]
    profiled_ms = timed_block()
    kern = {}
    for pl in plans:
        for kname, (nl, ms) in pl.profile_read(reset=True).items():
            a = kern.get(kname, (0, 0.0))
            kern[kname] = (a[0] + nl, a[1] + ms)
        pl.profile(False)
    st1 = [pl.stats() for pl in plans]
    lm.speculate = spec
    trials = lm.n_trials - trials0
    forwards = lm.n_forward - fwd0
    jacobians = lm.n_jacobian - jac0
    st = plan.stats()
    # profile evaluations really done in the profiled region: first pass and refinement queues (quad_level^2 nodes per
    # entry, every depth), value-only and derivative passes apart -- totals of every plan of the fit
    q2 = max([s.quad_level for s in plan.scene.sources] or [3]) ** 2
    spe = {"first": [0, 0], "queue": [0, 0]}
    for a, b in zip(st0, st1):
        for k in range(2):
            spe["first"][k] += b["cum_first_pass_evals"][k] - a["cum_first_pass_evals"][k]
            spe["queue"][k] += q2 * (sum(b["cum_queued"][k]) - sum(a["cum_queued"][k]))

    # ---- e2e: every step's data + weight come from pinned host memory and its result goes back to the host.
    #      Double-buffered like a survey pipeline would run it: while step k is fitted from one set of device buffers,
    #      step k+1's images upload into the other set on a copy stream (apb_plan_set_image_data rebinds the plan);
    #      every upload and every read-back lies inside the timed region.
    sets = [[dict(b) for b in plan.image_buffers], [{k: torch.empty_like(v) for k, v in b.items()} for b in plan.image_buffers]]
    pin = [{k: v.cpu().pin_memory() for k, v in bufs.items()} for bufs in plan.image_buffers]   # every local band / tile
    h2d = sum(v.numel() * 8 for pb in pin for v in pb.values())
    out_pin = torch.empty(len(x0) + 1, dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def upload(which):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[which])
            for bufs, pb in zip(sets[which], pin):
                for name, buf in bufs.items():
                    buf.copy_(pb[name], non_blocking=True)
            uploaded[which].record(copy_stream)

    def bind(which):
        for i, bufs in enumerate(sets[which]):
            if bufs:
                lm.set_image_data(i, bufs["data"], bufs.get("weight"))      # every plan of the fit

    def e2e_pass(n_steps):
        reset()
        barrier()
        for ev in consumed:
            ev.record()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for k in range(n_steps):
            flush.zero_()
            barrier()
            e0.record()
            cur = k % 2
            if k == 0:
                upload(cur)                      # nothing to hide the first upload behind
            torch.cuda.current_stream().wait_event(uploaded[cur])
            bind(cur)
            if k + 1 < n_steps:
                upload(1 - cur)                  # next step's images, concurrent with this step's fit
            one_iteration()
            consumed[cur].record()
            out_pin[:-1].copy_(lm.current_state, non_blocking=True)
            out_pin[-1] = lm.loss_history[-1]
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        bind(0)
        return total

    e2e_pass(max(min(warmup, 5), 3))     # untimed warm-up of the streaming path (copy stream, second buffer set, pinned staging)
    e2e_ms = allmax(e2e_pass(steps))
    m1 = clocks.mark()
    clock_summary = clocks.summary(m0, m1)
    e2e_value = units_per_step * steps / (e2e_ms * 1e-3)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": workload_text(wl, n_bands, world),
                   "l2": "256 MB buffer written between timed iterations (outside the event pairs)",
                   "timing": f"median of {len(blocks_ms)} block(s) of {steps} consecutive LM iterations, "
                             f"{sum(blocks_ms) * 1e-3:.2f} s timed in all; the iterations are those of consecutive fits of {FIT_ITERS} LM "
                             "iterations from the perturbed start (ended early at the noise floor of chi^2 or when LM cannot "
                             "improve it); the reference arm walks the same fits",
                   "kernel_timing": "one more block of the same K iterations with CUDA events around every launch "
                                    f"({profiled_ms / steps:.3f} ms/step with the event records)",
                   "params": len(x0), "lambda_trials_per_iter": trials / steps, "forwards_per_iter": forwards / steps,
                   "fit_restarts": state["restarts"],
                   "pcg_iterations_mean": (float(np.mean(lm.pcg_iterations)) if lm.pcg_iterations else None),
                   "pcg_solves": len(lm.pcg_iterations), "block_array_doubles": plan.block_doubles()},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_pin.numel() * 8,
                "pipeline": "double-buffered: step k+1's data + weight upload from pinned memory on a copy stream while step k "
                            "is fitted; the plan is rebound with apb_plan_set_image_data; uploads and read-back are inside the timed region"},
        "gpu_launches": launches,
        "blocks_ms": [round(b, 3) for b in blocks_ms[:32]],
    }
    if rank != 0:
        del lm, plan, plans, sets, pin
        torch.cuda.empty_cache()
        return None

    # ---- roofline of the dominant kernel (algorithmic work per DESIGN.md §4 / SURVEY.md §8d)
    dfma_tflops, copy_gbs = ctx["peaks"]
    hbm_peak, hbm_src = ctx["hbm_peak"], ctx["hbm_src"]
    work = algorithmic_work(plan.scene, spe, kern, n_fwd=forwards - jacobians, n_jac=jacobians, n_geo=trials)
    traffic_all = {}
    try:   # dram bytes per launch from the committed ncu --set full capture of this command
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass

    def roof_of(kname):
        nl, ms = kern[kname]
        w = work.get(kname)
        r = {"kernel": kname, "launches": nl, "avg_ms": ms / max(nl, 1),
             "share_of_step": ms / max(sum(v[1] for v in kern.values()), 1e-9)}
        tr = traffic_all.get(f"{wl}:{kname}", traffic_all.get(kname))
        if w is None:
            r.update({"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": tr})
        elif w["bound"] == "fp64":
            ach = w["flops"] / (ms * 1e-3) / 1e12
            r.update({"bound": "fp64", "achieved": ach, "peak": dfma_tflops, "unit": "TFLOP/s", "frac": ach / dfma_tflops,
                      "frac_nominal": w["flops_nominal"] / (ms * 1e-3) / 1e12 / dfma_tflops,
                      "traffic": tr, "peak_source": "apb_bench_peaks DFMA stream measured in this run "
                      "(MEASURED_PEAKS.json has no fp64 figure; nominal 37.2)", "algorithmic": w["what"]})
        else:
            ach = w["bytes"] / (ms * 1e-3) / 1e9
            r.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                      "traffic": tr, "peak_source": hbm_src, "algorithmic": w["what"],
                      "algorithmic_bytes_per_launch": w["bytes"] / max(nl, 1)})
        return r

    top = max(kern.items(), key=lambda kv: kv[1][1])[0] if kern else None
    roof = roof_of(top) if top else {"kernel": "none"}
    roof_all = {}
    for kname in kern:
        if kname in work and kern[kname][1] > 0:
            r = roof_of(kname)
            roof_all[kname] = {"bound": r["bound"], "frac": round(r["frac"], 4)}
            if "frac_nominal" in r:
                roof_all[kname]["frac_nominal"] = round(r["frac_nominal"], 4)
    out.update({
        "roofline": roof, "roofline_all": roof_all,
        "mpix_per_s_sampled": world * forwards * (n_pix_local / 1e6) / (profiled_ms * 1e-3),
        "spe_per_s": (sum(spe["first"]) + sum(spe["queue"])) / (profiled_ms * 1e-3),
        "kernel_ms": {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])},
        "refine_queue_last": st["queued"], "peaks_now": {"dfma_tflops": dfma_tflops, "copy_gbs": copy_gbs},
    })
    del lm, plan, plans, sets, pin
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import astrophot_b200 as ap
    from astrophot_b200 import cabi

    ap.AP_config.ap_device = f"cuda:{local}"
    if args.dtype == "f32":
        # the reference's switch (AP_config.py:7): images in fp32, profile kernels in fp32 arithmetic; the convolution
        # planes and the normal equations stay fp64
        ap.AP_config.ap_dtype = torch.float32
    dev = torch.device("cuda", local)
    main_wl = args.workload or default_workload(world)
    extras = []
    if args.workload is None and not args.no_extras:
        # the other BASELINE configurations at this world size, as extra keys of the line (shorter timed regions)
        extras = ["c4", "c2"] + (["c5s"] if world == 1 else [])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = None
    for key in ("hbm_gbs", "hbm_copy_gbs", "hbm_gbs_burst", "hbm_burst_gbs"):
        if isinstance(peaks.get(key), (int, float)):
            hbm_peak = float(peaks[key])
            break
    dfma_tflops, copy_gbs = cabi.bench_peaks()
    clocks = ClockSampler(local)
    clocks.__enter__()          # sampling spans warm-up, the timed region and the e2e region (all under load)
    ctx = {"world": world, "rank": rank, "dev": dev, "clocks": clocks, "peaks": (dfma_tflops, copy_gbs),
           "hbm_peak": hbm_peak or copy_gbs,
           "hbm_src": "MEASURED_PEAKS.json" if hbm_peak else "apb_bench_peaks fp64 copy measured in this run (MEASURED_PEAKS.json absent)",
           "flush": torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)}   # > 126 MB L2
    line = measure(main_wl, args, ctx, main=True)
    others = {}
    for wl in extras:
        try:
            r = measure(wl, args, ctx, main=False)
        except Exception as e:      # an extra workload must never cost the headline its line
            r = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0 and r is not None:
            keep = ("value", "unit", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "roofline", "roofline_all",
                    "kernel_ms", "error")
            others[wl] = {k: r[k] for k in keep if k in r}
    clocks.__exit__()
    if rank == 0:
        # ---- CPU baseline (bounded sample of the oracle port on the host cores, the GPU arm's iteration mix)
        cpu = None
        if world == 1 and not args.no_cpu:
            scene_c, x0_c = cpu_scene(main_wl)
            n_cpu = 4 if main_wl in ("c3", "c3s", "c3t") else 2
            times, _ = cpu_lm_iterations(scene_c, x0_c, 1, n_cpu, budget_s=40.0)
            fac, note, scale_info = cpu_scale(main_wl)
            cpu = dict({"value": fac / float(np.mean(times)), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": f"{len(times)} consecutive LM iterations of the numpy/scipy oracle port after 1 untimed one (FFT "
                                  "convolution; thread pool over sources and convolution planes + BLAS threads, ~1.2x over one "
                                  "thread), same workload" + note}, **scale_info)
        line["cpu_baseline"] = cpu
        if others:
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=20)
    ap_.add_argument("--warmup", type=int, default=5)
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--workload", default=None, choices=["c2", "c2b", "c3t", "c3s", "c3", "c4", "c5t", "c5s", "c5"],
                     help="default: c3 = BASELINE config[2], the 4096^2 crowded field (the configuration the 1-GPU target is "
                          "quoted on; N > 1: cut into N tiles), with c4 / c2 / c5s measured too and attached as "
                          "`other_workloads`.  c2 = config[1]; c3s / c3t = 1024^2 / 512^2 scale models of c3; c4 = config[3], "
                          "8-band joint fit on 2048^2 (strong scaling, 8/N bands per GPU); c5 = config[4], 16384^2 mosaic with "
                          "10000 galaxies, c5s / c5t its 2048^2 / 512^2 scale models")
    ap_.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                     help="f32: AP_config.ap_dtype = float32 -- the profile kernels (first pass, sub-pixel integration) "
                          "compute in single precision (parity bar 1e-5); default f64, the metric's precision")
    ap_.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap_.add_argument("--no-extras", action="store_true", help="only the headline workload")
    ap_.add_argument("--conv", default=None, choices=["direct", "fft"],
                     help="force one PSF-convolution kernel family (default: automatic, FFT for the 51x51 PSF)")
    args = ap_.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
