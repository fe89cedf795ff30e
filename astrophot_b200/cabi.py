"""ctypes binding of ``libastrophot_b200.so`` (C ABI in include/astrophot_b200.h).

This is the only door between the Python host code and the sm_100a kernels.
torch is used for device memory and the current stream, nothing else.  If the
library is missing or no CUDA device is present the product path raises
``NativeLibraryError`` — there is no CPU fallback by design.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import scene as sc
from .errors import NativeLibraryError

__all__ = ["lib", "load_library", "Plan", "lm_solve", "LIB_PATH"]

LIB_PATH = os.environ.get("APB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libastrophot_b200.so")

MAX_ELEM, MAX_PROF, MAX_DEPTH = sc.MAX_ELEM, sc.MAX_PROF, 4


class apb_param_t(C.Structure):
    _fields_ = [("transform", C.c_int32), ("_pad", C.c_int32), ("lo", C.c_double), ("hi", C.c_double)]


class apb_image_t(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("S", C.c_double * 4), ("rij", C.c_double * 2),
                ("rxy", C.c_double * 2), ("data", C.c_void_p), ("weight", C.c_void_p), ("mask", C.c_void_p),
                ("flags", C.c_int32), ("_pad", C.c_int32)]


class apb_psf_t(C.Structure):
    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("data", C.c_void_p), ("source", C.c_int32), ("_pad", C.c_int32)]


class apb_source_t(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("image", C.c_int32),
                ("out", C.c_int32 * 4), ("fwd", C.c_int32 * 4), ("jac", C.c_int32 * 4),
                ("n_elem", C.c_int32), ("slot", C.c_int32 * MAX_ELEM), ("cval", C.c_double * MAX_ELEM),
                ("n_prof", C.c_int32), ("prof", C.c_double * MAX_PROF),
                ("sampling_mode", C.c_int32), ("quad_init", C.c_int32), ("integrate_mode", C.c_int32),
                ("quad_level", C.c_int32), ("gridding", C.c_int32), ("max_depth", C.c_int32),
                ("ref_mode", C.c_int32), ("psf", C.c_int32), ("psf_shift", C.c_int32), ("conv_mode", C.c_int32),
                ("owner", C.c_int32), ("upscale", C.c_int32),
                ("tolerance", C.c_double), ("softening", C.c_double),
                ("mask", C.c_void_p), ("mask_rect", C.c_int32 * 4)]


class apb_owner_t(C.Structure):
    _fields_ = [("image", C.c_int32), ("out", C.c_int32 * 4), ("n_slot", C.c_int32), ("slot", C.c_int32 * MAX_ELEM)]


class apb_opts_t(C.Structure):
    _fields_ = [("queue_capacity", C.c_int64), ("flags", C.c_int32), ("n_owners", C.c_int32),
                ("owners", C.POINTER(apb_owner_t))]


class apb_stats_t(C.Structure):
    _fields_ = [("first_pass_evals", C.c_int64), ("queued", C.c_int64 * (MAX_DEPTH + 1)),
                ("launches", C.c_int64), ("overflow", C.c_int64),
                ("cum_passes", C.c_int64 * 2), ("cum_first_pass_evals", C.c_int64 * 2),
                ("cum_queued", (C.c_int64 * (MAX_DEPTH + 1)) * 2)]


class apb_kernel_time_t(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("total_ms", C.c_double)]


EXPORTS = ["apb_comm_alloc", "apb_comm_create", "apb_allreduce", "apb_comm_destroy", "apb_lm_trial_spec", "apb_plan_set_image_data", "apb_plan_block_doubles", "apb_plan_bind_blocks", "apb_lm_solve_sparse", "apb_lm_trial", "apb_lm_trial_begin", "apb_lm_trial_end", "apb_fft_length", "apb_plan_reserve", "apb_profile", "apb_profile_read", "apb_launch_count", "apb_bench_peaks", "apb_plan_create", "apb_plan_destroy", "apb_sample", "apb_jacobian", "apb_normal_eq", "apb_geodesic",
           "apb_chi2", "apb_lm_solve", "apb_plan_stats", "apb_last_error", "apb_version", "apb_chol_factor", "apb_chol_solve"]

_lib = None


def load_library(path=None):
    """dlopen the native library and declare its prototypes.  Loading does not
    need a GPU (the CPU test-suite checks the exports this way)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). astrophot_b200 has no CPU fallback.")
    try:
        L = C.CDLL(path)
    except OSError as e:
        raise NativeLibraryError(f"could not load {path}: {e}") from e
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    L.apb_plan_create.argtypes = [C.POINTER(apb_source_t), C.c_int, C.POINTER(apb_image_t), C.c_int,
                                  C.POINTER(apb_psf_t), C.c_int, C.POINTER(apb_param_t), C.c_int,
                                  C.POINTER(apb_opts_t), C.POINTER(C.c_void_p)]
    L.apb_plan_destroy.argtypes = [vp]
    L.apb_sample.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_void_p), vp]
    L.apb_jacobian.argtypes = [vp, dp, C.c_int, C.POINTER(C.c_void_p), vp]
    L.apb_normal_eq.argtypes = [vp, dp, C.c_int, dp, dp, dp, vp]
    L.apb_geodesic.argtypes = [vp, dp, dp, C.c_double, dp, vp]
    L.apb_chi2.argtypes = [vp, dp, dp, vp]
    L.apb_lm_solve.argtypes = [dp, dp, C.c_double, C.c_int, dp, ip, vp]
    L.apb_chol_factor.argtypes = [dp, C.c_double, C.c_int, dp, ip, vp]
    L.apb_chol_solve.argtypes = [dp, dp, C.c_int, dp, vp]
    L.apb_lm_solve_sparse.argtypes = [vp, dp, C.c_double, dp, dp, dp, C.c_double, C.c_int, vp]
    L.apb_plan_set_image_data.argtypes = [vp, C.c_int, dp, dp, dp, vp]
    L.apb_plan_block_doubles.argtypes = [vp]
    L.apb_plan_bind_blocks.argtypes = [vp, dp]
    L.apb_lm_trial.argtypes = [vp, vp, dp, dp, C.c_double, dp, C.c_double, C.c_double, dp, dp, dp, vp]
    L.apb_lm_trial_spec.argtypes = [vp, vp, vp, dp, dp, C.c_double, dp, C.c_double, C.c_double, dp, dp, dp, vp]
    L.apb_lm_trial_begin.argtypes = [vp, vp, dp, dp, C.c_double, dp, C.c_double, dp, dp, vp]
    L.apb_lm_trial_end.argtypes = [vp, dp, C.c_double, dp, dp, dp, dp, dp, vp]
    L.apb_plan_stats.argtypes = [vp, C.POINTER(apb_stats_t)]
    L.apb_plan_reserve.argtypes = [vp, C.POINTER(C.c_int64)]
    L.apb_profile.argtypes = [vp, C.c_int]
    L.apb_profile_read.argtypes = [vp, C.POINTER(apb_kernel_time_t), C.c_int, C.POINTER(C.c_int), C.c_int]
    L.apb_bench_peaks.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.apb_comm_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    L.apb_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.apb_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.apb_comm_destroy.argtypes = [C.c_void_p]
    L.apb_last_error.restype = C.c_char_p
    L.apb_version.restype = C.c_int
    for name in EXPORTS:
        if name not in ("apb_last_error",):
            getattr(L, name).restype = C.c_int if name != "apb_last_error" else C.c_char_p
    L.apb_last_error.restype = C.c_char_p
    L.apb_launch_count.restype = C.c_longlong
    L.apb_plan_block_doubles.restype = C.c_longlong
    _lib = L
    return L


def lib():
    return load_library()


def _check(rc, what):
    if rc != 0:
        msg = lib().apb_last_error()
        raise NativeLibraryError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def _require_cuda():
    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "astrophot_b200 evaluates models only on a CUDA device (B200, sm_100a); no CUDA device is visible "
            "and there is no CPU fallback.")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f64(x):
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    return t.to(device="cuda", dtype=torch.float64).contiguous()


class Plan:
    """A lowered model tree resident on the device (``apb_plan_t``)."""

    def __init__(self, scene, queue_capacity=0, conv=None, fused_integration=True, share=None, pooled_integration=False,
                 fp32=None):
        """``conv``: None (per-source psf_convolve_mode), "direct" or "fft" to force one
        convolution kernel family for every source (tests, benchmarks).
        ``share``: another Plan of the same scene whose device copies of data / weight / mask / PSFs
        are reused (the forward-only twin LM uses for the concurrent chi^2 pass).
        ``fp32``: profile kernels (first pass, mean reference, sub-pixel integration) in single-precision arithmetic;
        default: ``AP_config.ap_dtype == torch.float32``, the reference's switch (AP_config.py:7)."""
        _require_cuda()
        if fp32 is None:
            from . import AP_config
            fp32 = AP_config.ap_dtype == torch.float32
        self.fp32 = bool(fp32) and fused_integration
        L = lib()
        self.scene = scene
        self.n_par = scene.n_par
        self._keep = []          # tensors whose storage the plan points into
        self._masks = {}
        n_img, n_src, n_psf = len(scene.images), len(scene.sources), len(scene.psfs)
        imgs = (apb_image_t * max(n_img, 1))()
        self.shapes = []          # target images only (grids of auxiliary PSF models are no outputs)
        self._aux = []
        self.image_buffers = []   # per image: device tensors the plan reads (refill in place to stream new data)
        for i, im in enumerate(scene.images):
            imgs[i].H, imgs[i].W = im.H, im.W
            imgs[i].S[:] = [float(v) for v in np.asarray(im.S).reshape(4)]
            imgs[i].rij[:] = [float(v) for v in im.rij]
            imgs[i].rxy[:] = [float(v) for v in im.rxy]
            aux = bool(getattr(im, "aux", False))
            imgs[i].flags = 1 if aux else 0
            self._aux.append(aux)
            bufs = {}
            for name in ("data", "weight"):
                arr = getattr(im, name)
                if arr is not None:
                    t = share.image_buffers[i][name] if share is not None else _dev_f64(arr)
                    self._keep.append(t)
                    bufs[name] = t
                    setattr(imgs[i], name, t.data_ptr())
            self.image_buffers.append(bufs)
            if im.mask is not None:
                m = share._masks[i] if share is not None else \
                    torch.as_tensor(im.mask).to(device="cuda", dtype=torch.uint8).contiguous()
                self._keep.append(m)
                imgs[i].mask = m.data_ptr()
                self._masks[i] = m
            if not aux:
                self.shapes.append((im.H, im.W))
        psfs = (apb_psf_t * max(n_psf, 1))()
        self._psfs = []
        for i, ps in enumerate(scene.psfs):
            if getattr(ps, "source", -1) >= 0:      # stamp produced by a PSF-model source on every pass
                self._psfs.append(None)
                psfs[i].h, psfs[i].w, psfs[i].data, psfs[i].source = int(ps.shape[0]), int(ps.shape[1]), None, int(ps.source)
                continue
            t = share._psfs[i] if share is not None else _dev_f64(ps.data)
            self._keep.append(t)
            self._psfs.append(t)
            psfs[i].h, psfs[i].w, psfs[i].data, psfs[i].source = t.shape[0], t.shape[1], t.data_ptr(), -1
        pars = (apb_param_t * max(self.n_par, 1))()
        for k in range(self.n_par):
            pars[k].transform = int(scene.transform[k])
            pars[k].lo = 0.0 if np.isnan(scene.lo[k]) else float(scene.lo[k])
            pars[k].hi = 0.0 if np.isnan(scene.hi[k]) else float(scene.hi[k])
        srcs = (apb_source_t * max(n_src, 1))()
        for i, s in enumerate(scene.sources):
            c = srcs[i]
            c.kind, c.flags, c.image = s.kind, s.flags, s.image
            c.out[:], c.fwd[:], c.jac[:] = list(s.out), list(s.fwd), list(s.jac)
            c.n_elem = s.n_elem
            for e in range(s.n_elem):
                c.slot[e] = int(s.slot[e])
                c.cval[e] = float(s.cval[e])
            c.n_prof = len(s.prof)
            for k, v in enumerate(s.prof):
                c.prof[k] = float(v)
            c.sampling_mode, c.quad_init, c.integrate_mode = s.sampling_mode, s.quad_init, s.integrate_mode
            c.quad_level, c.gridding, c.max_depth = s.quad_level, s.gridding, s.max_depth
            c.ref_mode, c.psf, c.psf_shift = s.ref_mode, s.psf, s.psf_shift
            c.conv_mode = int(getattr(s, "conv_mode", 0))
            c.owner = int(getattr(s, "owner", -1))
            c.upscale = int(getattr(s, "upscale", 1) or 1)
            c.tolerance, c.softening = s.tolerance, s.softening
            mk = getattr(s, "mask", None)
            if mk is not None:
                key = id(mk)
                self._src_masks = getattr(self, "_src_masks", {})
                if key not in self._src_masks:      # (the pieces of a model cut into chunks or tiles share one mask)
                    t = torch.as_tensor(np.ascontiguousarray(mk)).to(device="cuda", dtype=torch.uint8).contiguous()
                    self._src_masks[key] = t
                    self._keep.append(t)
                t = self._src_masks[key]
                c.mask = t.data_ptr()
                c.mask_rect[:] = [int(s.mask_origin[0]), int(s.mask_origin[1]), int(t.shape[1]), int(t.shape[0])]
        opts = apb_opts_t(queue_capacity=int(queue_capacity), flags={None: 0, "auto": 0, "direct": 1, "fft": 2}[conv] | (0 if fused_integration else 4) | (8 if pooled_integration else 0) | (16 if self.fp32 else 0))
        owners = getattr(scene, "owners", None)
        if owners:       # models of the whole fit (the scene holds their pieces: lowering.tile_scene / shard_scene)
            otab = (apb_owner_t * len(owners))()
            for k, (img_id, rect, slots) in enumerate(owners):
                otab[k].image = int(img_id)
                otab[k].out[:] = [int(v) for v in rect]
                otab[k].n_slot = len(slots)
                for e, sl in enumerate(slots):
                    otab[k].slot[e] = int(sl)
            opts.n_owners = len(owners)
            opts.owners = otab
            self._keep.append(otab)
        handle = C.c_void_p()
        _check(L.apb_plan_create(srcs, n_src, imgs, n_img, psfs, n_psf, pars, self.n_par, C.byref(opts),
                                 C.byref(handle)), "apb_plan_create")
        self._h = handle
        self._L = L

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.apb_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- helpers -------------------------------------------------------------
    def _x(self, x):
        t = _dev_f64(x).reshape(-1)
        if t.numel() != self.n_par:
            raise NativeLibraryError(f"parameter vector has {t.numel()} elements, plan expects {self.n_par}")
        return t

    def _ptrs(self, tensors):
        """Pointer per plan image: the target images' output tensors in order, NULL for aux images."""
        arr = (C.c_void_p * max(len(self._aux), 1))()
        it = iter(tensors)
        for i, aux in enumerate(self._aux):
            arr[i] = None if aux else next(it).data_ptr()
        return arr

    # -- refinement-queue capacity ------------------------------------------------
    def reserve(self, caps=None):
        """Grow the sub-pixel refinement queues (after an overflow) and clear the flag."""
        arr = None
        if caps is not None:
            arr = (C.c_int64 * (MAX_DEPTH + 1))(*([0] + [int(c) for c in caps] + [0] * MAX_DEPTH)[: MAX_DEPTH + 1])
        _check(self._L.apb_plan_reserve(self._h, arr), "apb_plan_reserve")

    def _retry_on_overflow(self, call):
        """Run ``call`` (which launches on the device), synchronise, and if a refinement queue
        overflowed grow the queues from the observed counts and repeat."""
        for _ in range(12):
            out = call()
            if not self.stats()["overflow"]:
                return out
            self.reserve()
        raise NativeLibraryError("sub-pixel refinement queues keep overflowing; pass queue_capacity= explicitly")

    # -- entry points ------------------------------------------------------------
    def sample(self, x, as_rep=False):
        x = self._x(x)
        outs = [torch.empty(h, w, dtype=torch.float64, device="cuda") for h, w in self.shapes]
        self._retry_on_overflow(lambda: _check(
            self._L.apb_sample(self._h, x.data_ptr(), int(as_rep), self._ptrs(outs), _stream()), "apb_sample"))
        return outs

    def jacobian(self, x, as_rep=False):
        x = self._x(x)
        outs = [torch.empty(h, w, self.n_par, dtype=torch.float64, device="cuda") for h, w in self.shapes]
        self._retry_on_overflow(lambda: _check(
            self._L.apb_jacobian(self._h, x.data_ptr(), int(as_rep), self._ptrs(outs), _stream()), "apb_jacobian"))
        return outs

    def normal_eq(self, x, as_rep=True, out=None, check=False):
        """Returns (JtWJ (P,P), JtWr (P,), chi2 (2,): chi^2 and status flag) device tensors.
        Asynchronous; ``check=True`` synchronises and transparently repeats after a queue overflow
        (LM instead watches the sticky flag that comes back with every chi^2 record)."""
        if check:
            return self._retry_on_overflow(lambda: self.normal_eq(x, as_rep, out))
        x = self._x(x)
        P = self.n_par
        if out is None:
            out = (torch.empty(P, P, dtype=torch.float64, device="cuda"),
                   torch.empty(P, dtype=torch.float64, device="cuda"),
                   torch.empty(2, dtype=torch.float64, device="cuda"))
        H, g, c2 = out
        _check(self._L.apb_normal_eq(self._h, x.data_ptr(), int(as_rep), H.data_ptr() if H is not None else None,
                                     g.data_ptr(), c2.data_ptr(),
                                     _stream()), "apb_normal_eq")
        return H, g, c2

    def geodesic(self, xdh, h, d, out=None):
        xdh, h = self._x(xdh), self._x(h)
        if out is None:
            out = torch.empty(self.n_par, dtype=torch.float64, device="cuda")
        _check(self._L.apb_geodesic(self._h, xdh.data_ptr(), h.data_ptr(), float(d), out.data_ptr(), _stream()),
               "apb_geodesic")
        return out

    def lm_trial(self, H, g, L, x, d, acceleration, h_out, ha_out, rec, twin=None, donor=None):
        """One lambda-trial on the device (fit/lm.py:274-293); rec <- [chi2, flag, |a|, |h|].
        ``twin``: second Plan of the same scene for the concurrent chi^2 pass (acceleration == 0).
        ``donor``: the Plan holding the stamp Jacobian of the last normal_eq when this plan is a forward-only
        copy running a speculative trial (apb_lm_trial_spec)."""
        _check(self._L.apb_lm_trial_spec(self._h, twin._h if twin is not None else None,
                                         donor._h if donor is not None else None, H.data_ptr(), g.data_ptr(),
                                         float(L), x.data_ptr(), float(d),
                                         float(acceleration), h_out.data_ptr(), ha_out.data_ptr(), rec.data_ptr(), _stream()),
               "apb_lm_trial_spec")
        return rec

    def lm_trial_begin(self, H, g, L, x, d, h_out, buf, twin=None):
        """First half of a sharded trial: buf <- [local rpp (P), local chi2, #non-finite, #overflow]."""
        _check(self._L.apb_lm_trial_begin(self._h, twin._h if twin is not None else None, H.data_ptr(), g.data_ptr(),
                                          float(L), x.data_ptr(), float(d), h_out.data_ptr(), buf.data_ptr(), _stream()),
               "apb_lm_trial_begin")

    def lm_trial_end(self, H, L, x, h, buf, ha_out, rec):
        """Second half (after the all-reduce of buf): rec <- [chi2, flag, |a|, |h|], ha_out <- h."""
        _check(self._L.apb_lm_trial_end(self._h, H.data_ptr(), float(L), x.data_ptr(), h.data_ptr(), buf.data_ptr(),
                                        ha_out.data_ptr(), rec.data_ptr(), _stream()), "apb_lm_trial_end")
        return rec

    def solve_sparse(self, g, L, out=None, info=None, tol=0.0, max_iter=0, x0=None):
        """Damped LM solve by block-sparse PCG on the blocks of the last normal_eq (apb_lm_solve_sparse).
        Returns (h, info) device tensors, info = [iterations, |r|/|b|]; None if the plan cannot use it."""
        if out is None:
            out = torch.empty(self.n_par, dtype=torch.float64, device="cuda")
        if info is None:
            info = torch.zeros(2, dtype=torch.float64, device="cuda")
        if x0 is not None and x0.data_ptr() == out.data_ptr():
            x0 = x0.clone()
        rc = self._L.apb_lm_solve_sparse(self._h, g.data_ptr(), float(L), x0.data_ptr() if x0 is not None else None,
                                         out.data_ptr(), info.data_ptr(), float(tol), int(max_iter), _stream())
        if rc == 1:
            return None
        _check(rc, "apb_lm_solve_sparse")
        return out, info

    def set_image_data(self, i, data, weight=None, mask=None):
        """Rebind image ``i`` to other device tensors of the same shape (fp64 data / weight, uint8 mask); work enqueued on
        the current stream afterwards reads them.  ``image_buffers[i]`` follows."""
        for t in (data, weight):
            if t is not None and (t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or
                                  tuple(t.shape) != (self.scene.images[i].H, self.scene.images[i].W)):
                raise NativeLibraryError("set_image_data: contiguous fp64 CUDA tensors of the image's shape are required")
        if mask is not None:
            mask = mask.to(dtype=torch.uint8).contiguous()
        _check(self._L.apb_plan_set_image_data(self._h, int(i), data.data_ptr(), weight.data_ptr() if weight is not None else None,
                                               mask.data_ptr() if mask is not None else None, _stream()),
               "apb_plan_set_image_data")
        self._bound = getattr(self, "_bound", {})
        self._bound[i] = (data, weight, mask)          # keep the storage alive while the plan points at it
        self.image_buffers[i] = {k: v for k, v in (("data", data), ("weight", weight)) if v is not None}

    def block_doubles(self):
        """Length of the block-sparse J^T W J array (0: the plan has no block-sparse form)."""
        return int(self._L.apb_plan_block_doubles(self._h))

    def bind_blocks(self, extra=0):
        """Torch-owned device array that receives the block-sparse J^T W J of every normal_eq
        ([64 doubles per owner block | diag H]; same layout on every rank of a tile-sharded fit, so a sum
        all-reduce of it merges the ranks' normal equations).  None if the plan has no block-sparse form.
        ``extra``: doubles appended for the caller (J^T W r rides in the same exchange)."""
        n = int(self._L.apb_plan_block_doubles(self._h))
        if n <= 0:
            return None
        buf = torch.zeros(n + int(extra), dtype=torch.float64, device="cuda")
        _check(self._L.apb_plan_bind_blocks(self._h, buf.data_ptr()), "apb_plan_bind_blocks")
        self._keep.append(buf)
        return buf

    def chi2(self, x, out=None):
        """(sum W (Y - model)^2, finite flag) as a 2-element device tensor."""
        x = self._x(x)
        if out is None:
            out = torch.empty(2, dtype=torch.float64, device="cuda")
        _check(self._L.apb_chi2(self._h, x.data_ptr(), out.data_ptr(), _stream()), "apb_chi2")
        return out

    def profile(self, enable=True):
        _check(self._L.apb_profile(self._h, int(enable)), "apb_profile")

    def profile_read(self, reset=True):
        """{kernel name: (launches, total ms)} from CUDA events on the launching stream."""
        arr = (apb_kernel_time_t * 32)()
        n = C.c_int(0)
        _check(self._L.apb_profile_read(self._h, arr, 32, C.byref(n), int(reset)), "apb_profile_read")
        return {arr[i].name.decode(): (int(arr[i].launches), float(arr[i].total_ms)) for i in range(n.value)}

    def stats(self):
        st = apb_stats_t()
        _check(self._L.apb_plan_stats(self._h, C.byref(st)), "apb_plan_stats")
        return {"first_pass_evals": st.first_pass_evals, "queued": list(st.queued)[1:], "launches": st.launches,
                "overflow": st.overflow, "cum_passes": list(st.cum_passes),
                "cum_first_pass_evals": list(st.cum_first_pass_evals),
                "cum_queued": [list(st.cum_queued[k])[1:] for k in range(2)]}


class PeerComm:
    """Sum all-reduce over NVLink peer memory for the ranks of ONE node (``apb_allreduce``): one kernel on the current
    stream, deterministic rank-order sum.  ``group``: the torch.distributed group whose ranks take part (used once, to
    gather the cudaIpc handles).  Raises NativeLibraryError if the peers' memory cannot be mapped (ranks on different
    nodes, no P2P): callers fall back to the NCCL all-reduce."""

    def __init__(self, max_doubles, group=None):
        import torch.distributed as dist
        _require_cuda()
        L = lib()
        self._L, self._h = L, None
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.max_doubles = int(max_doubles)
        local = C.c_void_p()
        handle = (C.c_char * 64)()
        _check(L.apb_comm_alloc(self.max_doubles, C.byref(local), handle), "apb_comm_alloc")
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).cuda()
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
        h = C.c_void_p()
        rc = L.apb_comm_create(self.rank, self.world, local, blob, self.max_doubles, C.byref(h))
        # every rank must agree on whether the communicator exists
        ok = torch.tensor([1.0 if rc == 0 else 0.0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if rc != 0 or ok.item() == 0.0:
            msg = L.apb_last_error()
            if rc == 0:
                L.apb_comm_destroy(h)
            raise NativeLibraryError(f"peer-memory communicator unavailable: {msg.decode() if msg else '?'}")
        self._h = h

    def allreduce(self, t):
        """In-place sum of the contiguous fp64 CUDA tensor ``t`` over the ranks, on the current stream."""
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or t.numel() > self.max_doubles:
            raise NativeLibraryError("PeerComm.allreduce: contiguous fp64 CUDA tensor of at most max_doubles elements required")
        _check(self._L.apb_allreduce(self._h, t.data_ptr(), t.numel(), _stream()), "apb_allreduce")
        return t

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.apb_comm_destroy(h)
            except Exception:
                pass
            self._h = None


def lm_solve(H, g, L, out=None, info=None):
    """Damped LM step on the device (fit/lm.py:359-371)."""
    _require_cuda()
    P = g.numel()
    if out is None:
        out = torch.empty(P, dtype=torch.float64, device="cuda")
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _check(lib().apb_lm_solve(H.data_ptr(), g.data_ptr(), float(L), int(P), out.data_ptr(), info.data_ptr(),
                              _stream()), "apb_lm_solve")
    return out


def chol_factor(H, L, work=None, info=None):
    """Blocked Cholesky factor of the damped matrix of fit/lm.py:359-371 for systems beyond the single-CTA solver
    (apb_chol_factor).  Returns (work, info): the factor (P*P + 2 doubles) and a device int (0 = ok)."""
    _require_cuda()
    P = H.shape[0]
    if work is None or work.numel() < P * P + 2:
        work = torch.empty(P * P + 2, dtype=torch.float64, device=H.device)
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=H.device)
    H = H.contiguous()
    _check(lib().apb_chol_factor(H.data_ptr(), float(L), int(P), work.data_ptr(), info.data_ptr(), _stream()), "apb_chol_factor")
    return work, info


def chol_solve(work, rhs, out=None):
    """Solve with the factor of ``chol_factor`` (apb_chol_solve)."""
    _require_cuda()
    P = rhs.numel()
    rhs = rhs.contiguous()
    if out is None:
        out = torch.empty(P, dtype=torch.float64, device=rhs.device)
    _check(lib().apb_chol_solve(work.data_ptr(), rhs.data_ptr(), int(P), out.data_ptr(), _stream()), "apb_chol_solve")
    return out


def fft_length(n):
    """Transform length the FFT convolution picks for a padded stamp of n pixels."""
    return int(lib().apb_fft_length(int(n)))


def launch_count():
    return int(lib().apb_launch_count())


def bench_peaks():
    """(fp64 DFMA TFLOP/s, fp64 copy GB/s) measured now on the current device."""
    _require_cuda()
    a, b = C.c_double(0), C.c_double(0)
    _check(lib().apb_bench_peaks(C.byref(a), C.byref(b)), "apb_bench_peaks")
    return a.value, b.value
