#!/usr/bin/env python
"""Benchmark of the AstroPhot forward-model-and-fit hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Metric (BASELINE.json): LM iterations/s (model + PSF + J^T J), fp64.  A *step*
is one Levenberg-Marquardt iteration (fit/lm.py:450-467): one fused
sample+Jacobian+normal-equation build and k lambda-trials, each with a damped
solve, a geodesic pass and a chi^2 pass.  The fit restarts from the perturbed
start whenever it converges, so K steps are K real iterations drawn from the
fit trajectory.

Workload at N GPUs: an N-band joint fit (Target_Image_List; shared centre / q /
PA / n / Re, per-band Ie), one band per GPU, every band being BASELINE config[1]
— a PSF-convolved Sersic on 1024x1024 with a 51x51 Moffat PSF and threshold
sub-pixel integration.  At N=1 that is exactly config[1].  `value` counts
band-iterations per second (= LM iterations/s at N=1), weak scaling; the only
collective is the all-reduce of J^T W J / J^T W r / chi^2 (P^2+P+2 doubles).

`--impl reference` times the CPU oracle port of the reference algorithm
(oracle/astrophot_oracle.py, numpy + scipy FFT convolution like the reference's
default psf_convolve_mode) on the host cores, same config, one LM iteration
per step.  /root/reference is not needed at run time.
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


def build_joint(ap, n_bands, datas, size=SIZE):
    """n_bands PSF-convolved Sersic models sharing shape parameters
    (docs/source/tutorials/JointModels.ipynb recipe).  datas[b] = dict(data, variance) or None."""
    psf_np = ap.utils.moffat_psf(2.5, 3.0, PSF_W, 1.0)
    tars, models = [], []
    for b in range(n_bands):
        d = datas[b] if datas is not None else None
        kw = {} if d is None else {"variance": d["variance"]}
        tars.append(ap.image.Target_Image(data=np.zeros((size, size)) if d is None else d["data"], pixelscale=1.0,
                                          zeropoint=22.5, psf=ap.image.PSF_Image(data=psf_np, pixelscale=1.0), **kw))
    for b in range(n_bands):
        pars = band_truth(b)
        pars["center"] = [size / 2 + 0.3, size / 2 - 0.4]
        m = ap.models.AstroPhot_Model(name=f"band{b}", model_type="sersic galaxy model", target=tars[b],
                                      psf_mode="full", parameters=pars)
        if b > 0:
            for p in ("center", "q", "PA", "n", "Re"):
                m[p].value = models[0][p]
        models.append(m)
    if n_bands == 1:
        return models[0]
    return ap.models.AstroPhot_Model(name="joint", model_type="group model", models=models,
                                     target=ap.image.Target_Image_List(tars), psf_mode="full")


def make_data(truth, seed):
    rng = np.random.default_rng(seed)
    var = 0.1**2 + truth / 100.0
    return {"data": truth + rng.normal(size=truth.shape) * np.sqrt(var), "variance": var}


def start_state(x_rep, seed=2):
    rng = np.random.default_rng(1000 + seed)
    return np.asarray(x_rep, dtype=np.float64) + 0.05 * rng.normal(size=len(x_rep))


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
def cpu_lm_iteration_seconds(scene, x0, n_iter=1):
    import astrophot_oracle as orc

    t0 = time.perf_counter()
    res = orc.lm_fit(scene, x0, max_iter=n_iter, relative_tolerance=0.0, conv="fft")
    dt = time.perf_counter() - t0
    its = max(1, len(res["loss_history"]) - 1)
    return dt / its, res


def cpu_scene(n_bands=1, size=SIZE):
    """Scene tables for the CPU arm (host numpy only, no GPU needed)."""
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    from astrophot_b200.lowering import lower

    ap.AP_config.ap_device = "cpu"
    model = build_joint(ap, n_bands, None, size)
    scene, _ = lower(model)
    xv = model.parameters.vector_values().numpy()
    truth = orc.sample(scene, xv, as_rep=False, conv="fft")
    datas = [make_data(t, 10 + b) for b, t in enumerate(truth)]
    model = build_joint(ap, n_bands, datas, size)
    scene, _ = lower(model, for_fit=True)
    x0 = start_state(model.parameters.vector_representation().numpy())
    return scene, x0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    scene, x0 = cpu_scene(1)
    times = []
    for k in range(args.warmup + args.steps):
        dt, res = cpu_lm_iteration_seconds(scene, x0, 1)
        if k >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = 1e3 / ms
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"c2: 1 PSF-convolved Sersic, {SIZE}x{SIZE}, {PSF_W}x{PSF_W} Moffat PSF, threshold integration, LM fp64",
                   "note": "CPU oracle port of the reference algorithm (numpy + scipy FFT conv); one LM iteration from the perturbed start per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 full-size LM iteration (1 normal-equation build + lambda trials) per step"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# algorithmic work of the kernels on config[1] (one PSF-convolved Sersic), summed over the timed
# region: n_fwd value-only sampling passes and n_jac value+derivative passes
# ---------------------------------------------------------------------------
def _fft_len(n):
    from astrophot_b200 import cabi
    return cabi.fft_length(n)


def algorithmic_work(src, n_fwd, n_jac, n_geo=0):
    ow, oh = src.out[2], src.out[3]
    spw = PSF_W + 2                       # bilinear-shifted stamp keeps its 1-px pad
    b = (PSF_W + 2) // 2                  # psf_border_int = ceil((P+1)/2)
    ew, eh = ow + 2 * b, oh + 2 * b
    n_act = sum(1 for sl in src.slot if sl >= 0)
    nx, ny = _fft_len(ew), _fft_len(eh)
    nxh = nx // 2 + 1
    px = ow * oh
    # planes per pass: value-only: 1 image plane, 1 PSF plane, 1 product; derivative pass: value + (n_act-2)
    # non-centre planes in, 3 PSF planes (K, dK/dcx, dK/dcy), 1 + n_act products out
    in_f, k_f, j_f = 1, 1, 1
    in_j, k_j, j_j = 1 + (n_act - 2), 3, 1 + n_act
    tot = lambda f, j: n_fwd * f + n_jac * j
    rows = lambda n_in, n_k: n_in * (eh * ew * 8 + eh * nxh * 16) + n_k * (spw * spw * 8 + spw * nxh * 16)
    cols = lambda n_k, n_j: n_k * (spw * nxh * 16 + nxh * ny * 16) + n_j * (eh * nxh * 16 + nxh * ny * 16 + oh * nxh * 16)
    inv = lambda n_j: n_j * (oh * nxh * 16 + px * 8)
    planes = tot(j_f, j_j)
    return {
        "k_conv": {"bound": "fp64", "flops": 2.0 * spw * spw * px * planes,
                   "what": f"2*{spw}^2 flop x {px} px x {planes} planes"},
        "k_fft_rows": {"bound": "hbm", "bytes": float(tot(rows(in_f, k_f), rows(in_j, k_j))),
                       "what": f"real rows in (8 B/px) + half spectra out (16 B x {nxh}/row), {nx}-point rows"},
        "k_fft_cols": {"bound": "hbm", "bytes": float(tot(cols(k_f, j_f), cols(k_j, j_j))),
                       "what": f"per product: read {eh}x{nxh} spectrum + {nxh}x{ny} PSF spectrum, write {oh}x{nxh}; 16 B each"},
        "k_fft_rows_inv": {"bound": "hbm", "bytes": float(tot(inv(j_f), inv(j_j))),
                           "what": f"half spectra in (16 B x {nxh}/row) + real rows out (8 B/px)"},
        # normal equations: (n_act planes + weight + residual) x 8 B + mask per pixel, once per build
        "k_blocks": {"bound": "hbm", "bytes": float(n_jac * px * ((n_act + 2) * 8 + 1) + n_geo * px * ((n_act + 1) * 8 + 1)),
                     "what": f"J^T W J build: {n_act} derivative planes + weight + residual (8 B) + mask per pixel; "
                             f"geodesic J^T v: {n_act} planes + v + mask per pixel"},
    }


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
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
    n_bands = world
    dev = torch.device("cuda", local)

    # truth + noisy data for the band(s); every rank builds all bands' descriptions, data only for its own
    truth_model = build_joint(ap, 1, None)
    datas = []
    for b in range(n_bands):
        if b % world == rank:
            pars = band_truth(b)
            truth_model["Ie"].value = pars["Ie"]
            t = truth_model().data.cpu().numpy()
            datas.append(make_data(t, 10 + b))
        else:
            datas.append(None)
    model = build_joint(ap, n_bands, datas)
    x_true = model.parameters.vector_representation().numpy()
    x0 = start_state(x_true)
    lm = ap.fit.LM(model, initial_state=x0, max_iter=10**6, relative_tolerance=0.0, distributed=(world > 1),
                   conv=args.conv)
    plan = lm.plan
    n_pix_local = sum(h * w for h, w in plan.shapes)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > 126 MB L2

    state = {"fresh": True, "iters_in_fit": 0, "restarts": 0}

    def reset():
        lm.current_state = torch.as_tensor(x0, dtype=torch.float64, device=dev)
        lm.L = 1.0
        lm.loss_history = [lm._chi2_record(lm.current_state)]
        lm.L_history, lm.lambda_history = [lm.L], [x0.copy()]
        state["iters_in_fit"] = 0

    def one_iteration():
        """One pass of the `for iteration in range(max_iter)` loop of LM.fit."""
        from astrophot_b200.errors import OptimizeStop
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
        if len(lm.loss_history) >= 3 and abs(lm.loss_history[-3] - lm.loss_history[-1]) / lm.loss_history[-1] < 1e-9:
            state["restarts"] += 1
            reset()     # converged: start the next fit (outside the next step's timing)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.__enter__()          # sampling spans warm-up, the timed region and the e2e region (all under load)
    reset()
    m0 = clocks.mark()
    for _ in range(args.warmup):
        one_iteration()
    barrier()

    # ---- timed region (device-resident inputs): K iterations, L2 flushed between them
    def timed_steps():
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.zero_()
            barrier()
            ev[k][0].record()
            one_iteration()
            ev[k][1].record()
        barrier()
        return float(sum(a.elapsed_time(b) for a, b in ev))

    launches0 = cabi.launch_count()
    total_ms = timed_steps()
    launches = cabi.launch_count() - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_bands * args.steps / (total_ms * 1e-3)

    # ---- the same K iterations again with every kernel launch bracketed by CUDA events on its stream
    #      (per-kernel durations for the roofline; kept out of `value` because the event records cost time)
    reset()
    plan.profile(True)
    plan.profile_read(reset=True)
    trials0, fwd0, jac0 = lm.n_trials, lm.n_forward, lm.n_jacobian
    profiled_ms = timed_steps()
    kern = plan.profile_read(reset=True)
    plan.profile(False)
    trials = lm.n_trials - trials0
    forwards = lm.n_forward - fwd0
    jacobians = lm.n_jacobian - jac0
    st = plan.stats()

    # ---- e2e: every step streams the band's data + weight from pinned host memory and reads the result back
    pin = {k: v.cpu().pin_memory() for k, v in plan.image_buffers[0].items()}
    h2d = sum(v.numel() * 8 for v in pin.values())
    out_pin = torch.empty(len(x0) + 1, dtype=torch.float64).pin_memory()
    reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        barrier()
        e0.record()
        for name, buf in plan.image_buffers[0].items():
            buf.copy_(pin[name], non_blocking=True)
        one_iteration()
        out_pin[:-1].copy_(lm.current_state, non_blocking=True)
        out_pin[-1] = lm.loss_history[-1]
        e1.record()
        torch.cuda.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    m1 = clocks.mark()
    clocks.__exit__()
    clock_summary = clocks.summary(m0, m1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_bands * args.steps / (float(t.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic work per DESIGN.md §4 / SURVEY.md §8d)
    dfma_tflops, copy_gbs = cabi.bench_peaks()
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
    hbm_src = "MEASURED_PEAKS.json" if hbm_peak else "apb_bench_peaks fp64 copy measured in this run (MEASURED_PEAKS.json absent)"
    hbm_peak = hbm_peak or copy_gbs
    top = max(kern.items(), key=lambda kv: kv[1][1]) if kern else ("none", (0, 0.0))
    name, (n_launch, k_ms) = top
    work = algorithmic_work(plan.scene.sources[0], n_fwd=forwards - jacobians, n_jac=jacobians, n_geo=trials)
    roof = {"kernel": name, "launches": n_launch, "avg_ms": k_ms / max(n_launch, 1),
            "share_of_step": k_ms / max(sum(v[1] for v in kern.values()), 1e-9)}
    w = work.get(name)
    traffic = None
    try:   # dram bytes per launch from the committed ncu --set full capture of this command
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
    except Exception:
        pass
    if w is None:
        roof.update({"bound": "hbm", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": traffic})
    elif w["bound"] == "fp64":
        ach = w["flops"] / (k_ms * 1e-3) / 1e12
        roof.update({"bound": "fp64", "achieved": ach, "peak": dfma_tflops, "unit": "TFLOP/s", "frac": ach / dfma_tflops,
                     "traffic": traffic, "peak_source": "apb_bench_peaks DFMA stream measured in this run "
                     "(MEASURED_PEAKS.json has no fp64 figure; nominal 37.2)", "algorithmic": w["what"]})
    else:
        ach = w["bytes"] / (k_ms * 1e-3) / 1e9
        roof.update({"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                     "traffic": traffic, "peak_source": hbm_src, "algorithmic": w["what"],
                     "algorithmic_bytes_per_launch": w["bytes"] / max(n_launch, 1)})
    roof_all = {}
    for kname, (nl, ms) in kern.items():
        ww = work.get(kname)
        if ww is None or ms <= 0:
            continue
        if ww["bound"] == "fp64":
            roof_all[kname] = {"frac": ww["flops"] / (ms * 1e-3) / 1e12 / dfma_tflops, "bound": "fp64"}
        else:
            roof_all[kname] = {"frac": ww["bytes"] / (ms * 1e-3) / 1e9 / hbm_peak, "bound": "hbm"}

    # ---- CPU baseline (bounded sample: one full-size LM iteration of the oracle port)
    cpu = None
    if world == 1 and not args.no_cpu:
        scene_c, x0_c = cpu_scene(1)
        dt, _ = cpu_lm_iteration_seconds(scene_c, x0_c, 1)
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "1 full-size LM iteration of the numpy/scipy oracle port (FFT convolution), same workload"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"c2 x {n_bands} band(s): PSF-convolved Sersic, {SIZE}x{SIZE} per band, {PSF_W}x{PSF_W} Moffat PSF, "
                               "threshold sub-pixel integration, LM fp64, joint fit sharded 1 band/GPU",
                   "l2": "256 MB buffer written between timed iterations (outside the event pairs)",
                   "kernel_timing": "second pass of the same K iterations with CUDA events around every launch "
                                    f"({profiled_ms / args.steps:.3f} ms/step with the event records)",
                   "params": len(x0), "lambda_trials_per_iter": trials / args.steps, "forwards_per_iter": forwards / args.steps,
                   "fit_restarts": state["restarts"]},
        "clocks": clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_pin.numel() * 8},
        "gpu_launches": launches,
        "roofline": roof,
        "roofline_all": {k: {"bound": v["bound"], "frac": round(v["frac"], 4)} for k, v in roof_all.items()},
        "cpu_baseline": cpu,
        "mpix_per_s_sampled": n_bands * forwards * (n_pix_local / 1e6) / (profiled_ms * 1e-3),
        "kernel_ms": {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])},
        "refine_queue_last": st["queued"], "peaks_now": {"dfma_tflops": dfma_tflops, "copy_gbs": copy_gbs},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--gpus", type=int, default=1)
    ap_.add_argument("--steps", type=int, default=100)
    ap_.add_argument("--warmup", type=int, default=5)
    ap_.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap_.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap_.add_argument("--conv", default=None, choices=["direct", "fft"],
                     help="force one PSF-convolution kernel family (default: automatic, FFT for the 51x51 PSF)")
    args = ap_.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
