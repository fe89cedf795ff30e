"""Levenberg-Marquardt on the device (drop-in for ``ap.fit.LM``).

Public surface and control flow follow the reference (`fit/lm.py:15-539`,
`fit/base.py:14-156`): same constructor keywords, same damping schedule
(Lup x11 capped 1e9, Ldn /9 floored 1e-9), same geodesic-acceleration
curvature test, same convergence messages, same attributes
(``current_state``, ``loss_history``, ``L_history``, ``lambda_history``,
``message``, ``hess``, ``grad``, ``covariance_matrix``, ``res()`` ...).

What differs is where the work happens.  The reference builds a dense
(N_pix, P) Jacobian with forward-mode AD and multiplies it out with torch;
here one ``apb_normal_eq`` call samples the model, evaluates analytic
derivatives and accumulates J^T W J / J^T W r / chi^2 on the GPU without ever
forming J, ``apb_geodesic`` reuses the cached per-source stamp Jacobian for the
second-order term, and the damped solve stays on the device.  Per lambda-trial
the host reads back one 4-double record (chi^2, finite flag, |a|, |h|), which is
what the reference's ``.item()`` control flow needs.

Multi-GPU: pass ``process_group=`` (or have ``torch.distributed`` initialised and
pass ``distributed=True``).  Each rank owns the images (bands / tiles) its model
was built with; J^T W J, J^T W r, chi^2 and the geodesic right-hand side are
summed with an NCCL all-reduce, every rank solves the same small system.
"""
import numpy as np
import torch

from . import AP_config
from .errors import OptimizeStop
from .lowering import lower, shard_scene, tile_scene

__all__ = ["BaseOptimizer", "LM", "Iter", "Iter_LM"]


def _state_device():
    """Device of the optimiser's state vectors.  The native library launches on the CURRENT CUDA device and the plans
    allocate there (cabi), so the state lives there too: a multi-GPU run that calls ``torch.cuda.set_device(rank)`` but
    leaves ``AP_config.ap_device`` at its default would otherwise hand cuda:0 pointers to kernels on cuda:rank."""
    if torch.cuda.is_available() and str(AP_config.ap_device).startswith("cuda"):
        cur = torch.device("cuda", torch.cuda.current_device())
        want = torch.device(AP_config.ap_device)
        if want.index is not None and want.index != cur.index:
            AP_config.ap_logger.warning(
                f"AP_config.ap_device is {want} but the current CUDA device is {cur}: the fit runs on {cur} "
                "(set AP_config.ap_device to match, or call torch.cuda.set_device)")
        return cur
    return torch.device(AP_config.ap_device)


class BaseOptimizer:
    """State shared by optimisers (reference: `fit/base.py:26-129`)."""

    def __init__(self, model, initial_state=None, relative_tolerance=1e-3, fit_window=None, **kwargs):
        self.model = model
        self.verbose = kwargs.get("verbose", 0)
        self.fit_window = self.model.window if fit_window is None else (fit_window & self.model.window)
        if initial_state is None:
            self.model.initialize()
            initial_state = self.model.parameters.vector_representation()
        if isinstance(initial_state, torch.Tensor):
            initial_state = initial_state.detach().cpu().numpy()
        self.current_state = torch.as_tensor(np.asarray(initial_state, dtype=np.float64), dtype=torch.float64,
                                             device=_state_device())
        if self.verbose > 1:
            AP_config.ap_logger.info(f"initial state: {self.current_state}")
        self.max_iter = kwargs.get("max_iter", 100 * len(initial_state))
        self.iteration = 0
        self.save_steps = kwargs.get("save_steps", None)
        self.relative_tolerance = relative_tolerance
        self.lambda_history = []
        self.loss_history = []
        self.message = ""

    def fit(self):
        raise NotImplementedError("Please use a subclass of BaseOptimizer for optimization")

    def step(self, current_state=None):
        raise NotImplementedError("Please use a subclass of BaseOptimizer for optimization")

    def chi2min(self):
        return np.nanmin(self.loss_history)

    def res(self):
        ok = np.isfinite(self.loss_history)
        if np.sum(ok) == 0:
            AP_config.ap_logger.warning("Getting optimizer res with no real loss history, using current state")
            return self.current_state.detach().cpu().numpy()
        return np.array(self.lambda_history)[ok][np.argmin(np.array(self.loss_history)[ok])]

    def res_loss(self):
        ok = np.isfinite(self.loss_history)
        return np.min(np.array(self.loss_history)[ok])


class _QueueOverflow(Exception):
    """A sub-pixel refinement queue overflowed on the device: grow it and redo the work."""


class LM(BaseOptimizer):
    """Levenberg-Marquardt with geodesic-acceleration curvature control."""

    def __init__(self, model, initial_state=None, max_iter=100, relative_tolerance=1e-5, ndf=None, **kwargs):
        super().__init__(model, initial_state, max_iter=max_iter, relative_tolerance=relative_tolerance, **kwargs)
        from .cabi import Plan

        self.max_iter = max_iter
        self.max_step_iter = kwargs.get("max_step_iter", 10)
        self.curvature_limit = kwargs.get("curvature_limit", 1.0)
        self._Lup = kwargs.get("Lup", 11.0)
        self._Ldn = kwargs.get("Ldn", 9.0)
        self.L = kwargs.get("L0", 1.0)
        self.acceleration = kwargs.get("acceleration", 0.0)
        self.group = kwargs.get("process_group", None)
        self.distributed = self.group is not None or bool(kwargs.get("distributed", False))

        # lower the model once; Y, W and the mask (target mask | fit mask) go to the device inside the plan
        scene, info = lower(model, window=kwargs.get("fit_window", None), for_fit=True)
        if kwargs.get("W", None) is not None:
            W = torch.as_tensor(kwargs["W"], dtype=torch.float64).flatten()
            at = 0
            for im in scene.images:
                if im.aux:
                    continue
                im.weight = W[at : at + im.H * im.W].reshape(im.H, im.W)
                at += im.H * im.W
        # tiles=(ny, nx): cut every image into tiles (one big image sharded over the ranks);
        # the tiles are then dealt to the ranks like the bands of a joint fit
        if kwargs.get("tiles", None) is not None:
            scene = tile_scene(scene, int(kwargs["tiles"][0]), int(kwargs["tiles"][1]))
        if self.distributed and kwargs.get("shard_images", True):
            scene = shard_scene(scene, torch.distributed.get_rank(self.group), torch.distributed.get_world_size(self.group))
        self.scene, self.info = scene, info
        n_keep = 0
        for im in scene.images:
            if im.aux:
                continue
            n_keep += im.H * im.W if im.mask is None else int((~torch.as_tensor(im.mask).bool()).sum())
        if self.distributed:
            # the count of the WHOLE fit decides (lm.py:206-209): a rank whose tile is fully masked must not stop alone
            # while the others wait in this all-reduce
            t = torch.tensor([float(n_keep)], dtype=torch.float64, device=self.current_state.device)
            torch.distributed.all_reduce(t, group=self.group)
            n_keep = int(t.item())
        if n_keep == 0:
            raise OptimizeStop("No data to fit. All pixels are masked")
        self.plan = Plan(scene, conv=kwargs.get("conv", None), queue_capacity=kwargs.get("queue_capacity", 0))   # conv: force "direct"/"fft" (tests, benchmarks)
        self._covariance_matrix = None
        P = len(self.current_state)
        self.ndf = max(1.0, n_keep - P) if ndf is None else ndf
        dev = self.current_state.device
        self._small_solver_max = kwargs.get("small_solver_max", 159)   # single-CTA device solver up to this P
        self._sparse_solver = bool(kwargs.get("sparse_solver", True))
        # Large systems with a block-sparse form (crowded fields) never build the dense P x P matrix: the PCG works on
        # the <= 8x8 blocks inside the plan.  It is allocated (and that iteration's normal equations redone) only if
        # a dense fallback solve is ever needed.
        self._dense = not (self._sparse_solver and P > self._small_solver_max and self.plan.block_doubles() > 0
                           and P > kwargs.get("dense_hess_max", 4096))
        # J^T W J and J^T W r live in ONE buffer: a sharded fit sums them over the ranks with a single exchange
        self._Hg = torch.empty(P * P + P if self._dense else P, dtype=torch.float64, device=dev)
        self._H = self._Hg[: P * P].view(P, P) if self._dense else None
        self._g = self._Hg[-P:] if P > 0 else self._Hg
        self._c2 = torch.empty(2, dtype=torch.float64, device=dev)
        self._rpp = torch.empty(P, dtype=torch.float64, device=dev)
        self._rec = torch.empty(4, dtype=torch.float64, device=dev)
        self._h = torch.empty(P, dtype=torch.float64, device=dev)
        self._ha = torch.empty(P, dtype=torch.float64, device=dev)
        # one C-ABI call per lambda-trial when the system fits the single-CTA solver and no
        # collective sits between the pieces of a trial
        self._fused_trial = kwargs.get("fused_trial", True) and 0 < P <= min(159, kwargs.get("small_solver_max", 159)) and \
            (not self.distributed or self.acceleration == 0)
        self._tbuf = torch.empty(P + 3, dtype=torch.float64, device=dev)
        self._split_trial = bool(kwargs.get("split_trial", False)) and self.acceleration == 0
        # acceleration == 0 (default): chi2(x + h) is independent of the geodesic term; a forward-only twin
        # plan lets apb_lm_trial evaluate it concurrently with the geodesic pass
        self.plan2 = None
        # A model whose window exceeds image_chunksize is cut into one piece per Jacobian chunk in the main plan (the
        # reference's chunked Jacobian, lowering.lower).  Forward passes do not need the cut -- and pay for it, every
        # piece re-evaluates its PSF border: the lambda-trials then run on forward-only plans of the UNCUT scene and
        # read the stamp Jacobian of the main plan (apb_lm_trial_spec, donor).
        self.planF = None
        scene_f = scene
        if info.chunked and self._fused_trial and not self.distributed and not self._split_trial \
                and kwargs.get("tiles", None) is None and kwargs.get("W", None) is None:
            scene_f, _ = lower(model, window=kwargs.get("fit_window", None), for_fit=True, chunk_jacobian=False)
            self.planF = Plan(scene_f, conv=kwargs.get("conv", None), queue_capacity=kwargs.get("queue_capacity", 0),
                              share=self.plan)
        if self._fused_trial and self.acceleration == 0 and kwargs.get("overlap_trial", True):
            self.plan2 = Plan(scene_f, conv=kwargs.get("conv", None), queue_capacity=kwargs.get("queue_capacity", 0),
                              share=self.plan)
        # Speculative lambda search (single GPU, fused trial; OFF by default): the damping of the next trial is
        # L / Ldn after an improvement, so that trial can be evaluated on a second pair of forward-only plans and a
        # second stream while the current one runs (apb_lm_trial_spec reads the stamp Jacobian of the main plan).
        # Same trials, same decisions, same results.  Measured on B200 (config[1]): half of the consumed trials are
        # correct guesses, but two trials side by side slow each other by almost 2x (the persistent integration
        # kernel and the FFT passes already occupy every SM), and every wrong guess is wasted work: 654 -> 564 LM
        # iterations/s.  Kept as an option for problems small enough to leave SMs idle.
        self._lanes = None
        self._pending = {}
        self.n_spec_hits = self.n_spec_launched = 0
        if self.plan2 is not None and self.planF is None and not self.distributed and not self._split_trial \
                and kwargs.get("speculate", False):
            mk = lambda: torch.empty(P, dtype=torch.float64, device=dev)
            plan3 = Plan(scene, conv=kwargs.get("conv", None), queue_capacity=kwargs.get("queue_capacity", 0), share=self.plan)
            plan4 = Plan(scene, conv=kwargs.get("conv", None), queue_capacity=kwargs.get("queue_capacity", 0), share=self.plan)
            self._spec_stream = torch.cuda.Stream(device=dev)
            self._lanes = [
                {"plan": self.plan, "twin": self.plan2, "donor": None, "h": self._h, "ha": self._ha, "rec": self._rec,
                 "event": torch.cuda.Event(), "busy": False},
                {"plan": plan3, "twin": plan4, "donor": self.plan, "h": mk(), "ha": mk(),
                 "rec": torch.empty(4, dtype=torch.float64, device=dev), "event": torch.cuda.Event(), "busy": False}]
            self._ev_ready = torch.cuda.Event()
        self.speculate = self._lanes is not None      # may be switched off between steps (per-kernel timing)
        self.hess = self.grad = None
        self._hess_version, self._factor_key, self._factor = 0, None, None
        self._blocks_version = -1       # _hess_version whose blocks the plan holds (set by _step, not by natural-units builds)
        # sharded fit of a large system: the ranks merge their normal equations as the block-sparse J^T W J
        # (a few MB, same owner layout on every rank) instead of the dense P x P matrix
        self._blk = None
        self._hess_reduced = True
        if self.distributed and self._sparse_solver and P > self._small_solver_max:
            # [blocks | diag H | J^T W r] in one array: one exchange per normal-equation build
            self._blkg = self.plan.bind_blocks(extra=P)
            if self._blkg is not None:
                self._blk = self._blkg[:-P]
                self._g = self._blkg[-P:]
                self._hs = torch.empty(P + 2, dtype=torch.float64, device=dev)
        # the exchange itself: one kernel over NVLink peer memory (apb_allreduce) when every rank of the group can map
        # the others' memory (one node), NCCL otherwise
        self._peer = None
        if self.distributed and kwargs.get("peer_allreduce", True) and torch.cuda.is_available():
            import os
            need = max(self._Hg.numel(), (self._blkg.numel() if self._blk is not None else 0), P + 3, 8)
            if os.environ.get("APB_NO_PEER", "0") in ("", "0") and need * 8 <= (256 << 20):
                try:
                    from .cabi import PeerComm
                    self._peer = PeerComm(need, group=self.group)
                except Exception as e:      # ranks on several nodes, no P2P: NCCL does it
                    if self.verbose > 0:
                        AP_config.ap_logger.info(f"peer-memory all-reduce unavailable ({e}); using NCCL")
        self.pcg_iterations = []
        self._warm = None               # (hess version, h, solve(rpp)) of the previous lambda-trial
        self._pcg_tol = float(kwargs.get("pcg_tol", 1e-12))   # relative residual of the damped solve (accepted up to 1e-10)
        self.n_forward = self.n_jacobian = self.n_trials = 0

    # -- damping -------------------------------------------------------------
    def Lup(self):
        self.L = min(1e9, self.L * self._Lup)

    def Ldn(self):
        self.L = max(1e-9, self.L / self._Ldn)

    # -- device pieces ----------------------------------------------------------
    def _allreduce(self, t):
        if self.distributed:
            if self._peer is not None and t.is_contiguous() and t.numel() <= self._peer.max_doubles:
                self._peer.allreduce(t)
            else:
                torch.distributed.all_reduce(t, group=self.group)
        return t

    def _chi2_record(self, x):
        """chi^2/ndf (host float) of the model at x."""
        self.n_forward += 1
        for _ in range(12):
            out = self._allreduce_chi((self.planF or self.plan).chi2(x, out=self._c2))
            c, ok = out.tolist()
            if ok >= 0.0:
                return c / self.ndf if ok >= 1.0 else float("nan")
            for pl in self.all_plans:
                pl.reserve()
        raise OptimizeStop("sub-pixel refinement queues keep overflowing")

    def _allreduce_chi(self, c2):
        if self.distributed:
            # per-rank flag: 1 ok, 0 non-finite, -1 queue overflow -> summed as (bad, overflow) counts
            rec = torch.stack([c2[0], (c2[1] == 0).to(c2.dtype), (c2[1] < 0).to(c2.dtype)])
            self._allreduce(rec)
            c2[0] = rec[0]
            c2[1] = torch.where(rec[2] > 0, -1.0, (rec[1] == 0).to(c2.dtype))
        return c2

    def set_image_data(self, i, data, weight=None, mask=None):
        """Point image ``i`` of the fit at other device tensors of the same shape (the next exposure of a survey, uploaded
        while the current one was fitted).  EVERY plan of the optimiser is rebound -- the main plan builds the normal
        equations, the forward-only twins evaluate the trial chi^2: rebinding only one of them would compare chi^2 of
        different data and take wrong accept / reject decisions without any error."""
        for pl in self.all_plans:
            pl.set_image_data(i, data, weight, mask if mask is not None else pl._masks.get(i))
        self._covariance_matrix = None

    @property
    def all_plans(self):
        """Every plan this optimiser launches work on (main, chi^2 twin, speculative pair)."""
        out = [self.plan] + ([self.plan2] if self.plan2 is not None else []) + \
            ([self.planF] if self.planF is not None else [])
        if self._lanes is not None:
            out += [self._lanes[1]["plan"], self._lanes[1]["twin"]]
        return out

    def _launch_trial(self, k, L, x, d):
        """Enqueue the trial at damping L on lane k (0: main stream and plans, 1: speculative stream and plans)."""
        ln = self._lanes[k]
        main = torch.cuda.current_stream()
        stream = main if k == 0 else self._spec_stream
        if k == 1:
            stream.wait_event(self._ev_ready)       # the normal equations of this iteration (main stream)
            x.record_stream(stream)
        with torch.cuda.stream(stream):
            ln["plan"].lm_trial(self.hess, self.grad, L, x, d, self.acceleration, ln["h"], ln["ha"], ln["rec"],
                                twin=ln["twin"], donor=ln["donor"])
            ln["event"].record(stream)
        ln["busy"] = True
        self.n_forward += 2
        self.n_spec_launched += 1

    def _trial_result(self, x, d):
        """(chi2 sum, flag, |a|, |h|, ha) of the trial at the current damping: taken from the lane that already runs
        it, else launched now; the likely next trial (L / Ldn) is started on the other lane before waiting."""
        k = self._pending.pop(self.L, None)
        if k is None:
            k = 0 if not any(v == 0 for v in self._pending.values()) else 1
            for Lp in [Lp for Lp, v in self._pending.items() if v == k]:
                del self._pending[Lp]              # a stale guess on that lane: it simply runs out
            self._launch_trial(k, self.L, x, d)
        else:
            self.n_spec_hits += 1
        nxt = max(1e-9, self.L / self._Ldn)
        other = 1 - k
        if nxt != self.L and nxt not in self._pending:
            for Lp in [Lp for Lp, v in self._pending.items() if v == other]:
                del self._pending[Lp]
            self._launch_trial(other, nxt, x, d)
            self._pending[nxt] = other
        ln = self._lanes[k]
        ln["event"].synchronize()
        ln["busy"] = False
        csum, ok, na, nh = ln["rec"].tolist()
        return csum, ok, na, nh, ln["ha"].clone()

    def _solve(self, L, rhs, loose=False, x0=None):
        from .cabi import lm_solve

        P = rhs.numel()
        if P <= self._small_solver_max:
            return lm_solve(self.hess, rhs, L)
        # large systems, no parameter shared between sources: block-sparse PCG on the <= 8x8 source-pair blocks the
        # normal-equation kernels produced (one cooperative launch); checked by its final relative residual
        if self._sparse_solver and (not self.distributed or self._blk is not None) \
                and self._blocks_version == self._hess_version:
            # `loose`: the geodesic correction a = -solve(rpp)/2 when acceleration == 0 only enters the trial through
            # the curvature ratio |a| / |h| compared with curvature_limit (lm.py:282-306): 1e-8 is plenty
            tol = 1e-8 if loose else self._pcg_tol
            if self.distributed:
                # every rank solves the same merged system (identical bits in: rank-order all-reduce; deterministic
                # solver); rank 0's answer is still the one all use -- a guard that costs one P-double exchange
                res = self.plan.solve_sparse(rhs.contiguous(), L, out=self._hs[:P], info=self._hs[P:], tol=tol, x0=x0)
                if self._peer is not None:
                    if torch.distributed.get_rank(self.group) != 0:
                        self._hs.zero_()
                    self._peer.allreduce(self._hs)          # = broadcast of rank 0's answer
                else:
                    torch.distributed.broadcast(self._hs, src=torch.distributed.get_global_rank(self.group, 0)
                                                if self.group is not None else 0, group=self.group)
                res = (self._hs[:P].clone(), self._hs[P:])
            else:
                res = self.plan.solve_sparse(rhs.contiguous(), L, tol=tol, x0=x0)
            if res is None:
                self._sparse_solver = False
            else:
                h, info = res
                its, rel = info.tolist()
                self.pcg_iterations.append(int(its))
                if rel <= (1e-7 if loose else 1e-10):
                    return h
        if self._H is None:
            # first dense fallback of a fit that ran on the blocks alone: build the dense matrix of this iteration
            self._H = torch.empty(P, P, dtype=torch.float64, device=rhs.device)
            g = torch.empty_like(self._g)
            self.plan.normal_eq(self._x_hess, as_rep=True, out=(self._H, g, torch.empty_like(self._c2)))
            if self._blk is not None:
                self._allreduce(self._blk)     # the rebuild overwrote the merged blocks with this rank's
            self.hess = self._H
            self._hess_reduced = not self.distributed
        if self.distributed and not self._hess_reduced:
            self._allreduce(self._H)
            self._hess_reduced = True
        # otherwise: same damped matrix (lm.py:359-371), dense.  It is symmetric positive definite for L > 0, so it is
        # Cholesky-factored once per (H, L) -- apb_chol_factor, a blocked factorisation in one cooperative kernel -- and
        # the factor serves both solves of a lambda-trial (h and the geodesic correction).  A matrix the factorisation
        # rejects (a non-finite or non-positive pivot: the trial is lost anyway) goes through LU like the reference's.
        from .cabi import chol_factor, chol_solve
        key = (self._hess_version, float(L))
        if self._factor_key != key:
            self._chol_work, info = chol_factor(self.hess, L, work=getattr(self, "_chol_work", None))
            if int(info.item()) == 0:
                self._factor = ("chol", self._chol_work)
            else:
                A = self.hess / (1.0 + L)
                d = torch.diagonal(self.hess)
                A.diagonal().copy_(d + L * (1.0 + d))
                self._factor = ("lu",) + tuple(torch.linalg.lu_factor(A))
                del A
            self._factor_key = key
        if self._factor[0] == "chol":
            return chol_solve(self._factor[1], rhs)
        return torch.linalg.lu_solve(self._factor[1], self._factor[2], rhs.reshape(-1, 1)).reshape(-1)

    @torch.no_grad()
    def step(self, chi2):
        """One LM iteration (reference: `fit/lm.py:248-357`).  If the device reports a
        refinement-queue overflow the queues are grown and the iteration is redone from its
        (unchanged) starting state."""
        L0 = self.L
        for _ in range(12):
            try:
                return self._step(chi2)
            except _QueueOverflow:
                for pl in self.all_plans:
                    pl.reserve()
                self._pending.clear()
                self.L = L0
        raise OptimizeStop("sub-pixel refinement queues keep overflowing")

    def _step(self, chi2):
        """Normal equations once, then search over the damping parameter."""
        x = self.current_state
        self._x_hess = x
        if self._lanes is not None:
            # a guess of the previous iteration may still be running: it reads H, g and the stamp Jacobian
            self._pending.clear()
            torch.cuda.current_stream().wait_event(self._lanes[1]["event"])
        self.plan.normal_eq(x, as_rep=True, out=(self._H, self._g, self._c2))
        if self._lanes is not None:
            self._ev_ready.record()
        self.n_forward += 1
        self.n_jacobian += 1
        if self.distributed:
            if self._blk is not None:
                self._allreduce(self._blkg)    # [blocks | diag H | g]; the dense copy stays local until a dense fallback asks for it
                self._hess_reduced = False
            else:
                self._allreduce(self._Hg)      # [H | g] (g alone when the dense matrix is not kept)
        self.hess, self.grad = self._H, self._g
        self._hess_version += 1
        self._blocks_version = self._hess_version
        init_chi2 = chi2
        nostep = True
        best = (torch.zeros_like(x), init_chi2, self.L)
        scary = (None, init_chi2, self.L)
        direction = "none"
        d = 0.1
        for it in range(self.max_step_iter):
            self.n_trials += 1
            if it > self.max_step_iter / 2 and self.L < 1e-3:
                self.L = 1.0
            if self._fused_trial:
                self.n_forward += 2
                if self.distributed or self._split_trial:
                    # sharded pixels: local rpp / chi2, one all-reduce of P + 3 doubles, then the second solve
                    self.plan.lm_trial_begin(self.hess, self.grad, self.L, x, d, self._h, self._tbuf, twin=self.plan2)
                    self._allreduce(self._tbuf)
                    self.plan.lm_trial_end(self.hess, self.L, x, self._h, self._tbuf, self._ha, self._rec)
                elif self._lanes is not None and self.speculate:
                    self.n_forward -= 2                        # counted per launch (guesses included)
                    csum, ok, na, nh, ha = self._trial_result(x, d)
                    if ok < 0.0:
                        raise _QueueOverflow()
                    chi2 = csum / self.ndf if ok >= 1.0 else float("nan")
                elif self.planF is not None:
                    self.planF.lm_trial(self.hess, self.grad, self.L, x, d, self.acceleration, self._h, self._ha,
                                        self._rec, twin=self.plan2, donor=self.plan)
                else:
                    self.plan.lm_trial(self.hess, self.grad, self.L, x, d, self.acceleration, self._h, self._ha,
                                       self._rec, twin=self.plan2)
                if not (self._lanes is not None and self.speculate):
                    csum, ok, na, nh = self._rec.tolist()          # the one host sync of this trial
                    if ok < 0.0:
                        raise _QueueOverflow()
                    ha = self._ha.clone()
                chi2 = csum / self.ndf if ok >= 1.0 else float("nan")
            else:
                ha, chi2, na, nh = self._trial_pieces(x, d)
            if self.verbose > 1:
                AP_config.ap_logger.info(f"sub step L: {self.L}, Chi^2/DoF: {chi2}")
            if not np.isfinite(chi2):
                if self.verbose > 1:
                    AP_config.ap_logger.info("Skip due to non-finite values")
                self.Lup()
                if direction == "better":
                    break
                direction = "worse"
                continue
            if chi2 <= scary[1]:
                scary = (ha, chi2, self.L)
            rho = na / nh if nh > 0 else float("nan")
            if rho > self.curvature_limit:
                if self.verbose > 1:
                    AP_config.ap_logger.info("Skip due to large curvature")
                self.Lup()
                if direction == "better":
                    break
                direction = "worse"
                continue
            if chi2 < best[1]:
                if self.verbose > 1:
                    AP_config.ap_logger.info("new best chi^2")
                best = (ha, chi2, self.L)
                nostep = False
                self.Ldn()
                if self.L <= 1e-8 or direction == "worse":
                    break
                direction = "better"
            elif chi2 > best[1] and direction in ("none", "worse"):
                if self.verbose > 1:
                    AP_config.ap_logger.info("chi^2 is worse")
                self.Lup()
                if self.L == 1e9:
                    break
                direction = "worse"
            else:
                break
            if (best[1] - init_chi2) / init_chi2 < -0.1:
                if self.verbose > 1:
                    AP_config.ap_logger.info("Large step taken, ending search for good step")
                break
        if nostep:
            if scary[0] is not None:
                if self.verbose > 1:
                    AP_config.ap_logger.warning("no low curvature step found, taking high curvature step")
                return scary
            raise OptimizeStop("Could not find step to improve chi^2")
        return best

    def _trial_pieces(self, x, d):
        """One lambda-trial from separate calls (distributed fits need an all-reduce between the
        pieces; large systems use the library solver).  Returns (ha, chi2/ndf, |a|, |h|)."""
        # the previous trial of this LM iteration solved the same matrix with another damping: its solutions are the
        # starting points of this trial's two PCG solves
        warm = self._warm if self._warm is not None and self._warm[0] == self._hess_version else (None, None, None)
        h = self._solve(self.L, self.grad, x0=warm[1])
        # geodesic acceleration (second directional derivative along h)
        self.n_forward += 1
        rpp = self._allreduce(self.plan.geodesic(x + d * h, h, d, out=self._rpp))
        if self.L > 1e-4:
            s2 = self._solve(self.L, rpp, loose=(self.acceleration == 0), x0=warm[2])
            a = -s2 / 2
        else:
            s2, a = None, torch.zeros_like(h)
        self._warm = (self._hess_version, h.clone(), None if s2 is None else s2.clone())
        ha = h + a * self.acceleration
        self.n_forward += 1
        c2 = self._allreduce_chi(self.plan.chi2(x + ha, out=self._c2))
        self._rec[0:2] = c2
        self._rec[2] = torch.linalg.norm(a)
        self._rec[3] = torch.linalg.norm(h)
        csum, ok, na, nh = self._rec.tolist()          # the one host sync of this trial
        if ok < 0.0:
            raise _QueueOverflow()
        chi2 = csum / self.ndf if ok >= 1.0 else float("nan")
        return ha, chi2, na, nh

    @torch.no_grad()
    def fit(self):
        """Iterate ``step`` to convergence (reference: `fit/lm.py:428-493`)."""
        if len(self.current_state) == 0:
            if self.verbose > 0:
                AP_config.ap_logger.warning("No parameters to optimize. Exiting fit")
            return self
        self._covariance_matrix = None
        self.loss_history = [self._chi2_record(self.current_state)]
        self.L_history = [self.L]
        self.lambda_history = [self.current_state.detach().cpu().numpy().copy()]
        for iteration in range(self.max_iter):
            if self.verbose > 0:
                AP_config.ap_logger.info(f"Chi^2/DoF: {self.loss_history[-1]}, L: {self.L}")
            try:
                res = self.step(chi2=self.loss_history[-1])
            except OptimizeStop:
                if self.verbose > 0:
                    AP_config.ap_logger.warning("Could not find step to improve Chi^2, stopping")
                self.message = self.message + "fail. Could not find step to improve Chi^2"
                break
            self.L = res[2]
            self.current_state = (self.current_state + res[0]).detach()
            self.L_history.append(self.L)
            self.loss_history.append(res[1])
            self.lambda_history.append(self.current_state.detach().cpu().numpy().copy())
            self.iteration += 1
            self.Ldn()
            if len(self.loss_history) >= 3:
                if (self.loss_history[-3] - self.loss_history[-1]) / self.loss_history[-1] < self.relative_tolerance \
                        and self.L < 0.1:
                    self.message = self.message + "success"
                    break
            if len(self.loss_history) > 10:
                if (self.loss_history[-10] - self.loss_history[-1]) / self.loss_history[-1] < self.relative_tolerance:
                    self.message = self.message + "success by immobility. Convergence not guaranteed"
                    break
        else:
            self.message = self.message + "fail. Maximum iterations"
        if self.verbose > 0:
            AP_config.ap_logger.info(
                f"Final Chi^2/DoF: {self.loss_history[-1]}, L: {self.L_history[-1]}. Converged: {self.message}")
        self.model.parameters.vector_set_representation(self.res())
        return self

    # -- uncertainties (natural parameters; reference: lm.py:408-425,495-539) -----
    @torch.no_grad()
    def update_hess_grad(self, natural=False):
        if natural:
            xv = self.model.parameters.vector_transform_rep_to_val(self.current_state.detach().cpu())
            H, g, _ = self.plan.normal_eq(xv, as_rep=False, check=True)
        else:
            H, g, _ = self.plan.normal_eq(self.current_state, as_rep=True, check=True)
        if self.distributed:
            self._allreduce(H)
            self._allreduce(g)
        self.hess, self.grad = H, g
        self._hess_version += 1

    @property
    @torch.no_grad()
    def covariance_matrix(self):
        if self._covariance_matrix is not None:
            return self._covariance_matrix
        self.update_hess_grad(natural=True)
        try:
            self._covariance_matrix = torch.linalg.inv(self.hess)
        except Exception:
            AP_config.ap_logger.warning(
                "WARNING: Hessian is singular, likely at least one model is non-physical. Will massage Hessian to "
                "continue but results should be inspected.")
            self.hess += torch.eye(len(self.grad), dtype=self.hess.dtype, device=self.hess.device) * (
                torch.diag(self.hess) == 0)
            self._covariance_matrix = torch.linalg.inv(self.hess)
        return self._covariance_matrix

    @torch.no_grad()
    def update_uncertainty(self):
        cov = self.covariance_matrix
        if torch.all(torch.isfinite(cov)):
            try:
                self.model.parameters.vector_set_uncertainty(torch.sqrt(torch.abs(torch.diag(cov))).cpu())
            except RuntimeError as e:
                AP_config.ap_logger.warning(f"Unable to update uncertainty due to: {e}")
        else:
            AP_config.ap_logger.warning("Unable to update uncertainty due to non finite covariance matrix")


class Iter(BaseOptimizer):
    """Fit the sub-models of a group one at a time on the residual image, over and over (drop-in for
    ``ap.fit.Iter``, reference `fit/iterative.py:19-180`): each sub-step subtracts every *other* model from the
    target, fits the one model to what is left with ``method`` (default: the device LM above), and puts it back.
    The control flow, convergence test and attributes follow the reference; the model images and the sub-fits run
    on the GPU (every sub-fit is one plan: its sampling, normal equations and damped solves stay on the device)."""

    def __init__(self, model, method=LM, initial_state=None, max_iter=100, method_kwargs=None, **kwargs):
        super().__init__(model, initial_state, max_iter=max_iter, **kwargs)
        self.max_iter = max_iter
        self.method = method
        self.method_kwargs = dict(method_kwargs or {})
        sub = self.model.target[self.model.window]
        self.ndf = sub.flatten("data").numel() - len(self.current_state)
        if self.model.target.has_mask:
            self.ndf -= int(torch.sum(sub.flatten("mask")).item())
        self._count_finish = 0

    def sub_step(self, model):
        """One model against the residual of all the others (reference: `fit/iterative.py:67-82`)."""
        self.Y -= model()
        initial_target = model.target
        model.target = model.target[model.window] - self.Y[model.window]
        res = self.method(model, **self.method_kwargs).fit()
        self.Y += model()
        if self.verbose > 1:
            AP_config.ap_logger.info(res.message)
        model.target = initial_target

    @torch.no_grad()
    def step(self):
        """One sweep over the sub-models, then chi^2 of the whole model (reference: `fit/iterative.py:84-135`)."""
        if self.verbose > 0:
            AP_config.ap_logger.info("--------iter-------")
        for model in self.model.models.values():
            if self.verbose > 0:
                AP_config.ap_logger.info(model.name)
            self.sub_step(model)
        self.current_state = torch.as_tensor(self.model.parameters.vector_representation().numpy(), dtype=torch.float64,
                                             device=_state_device())
        self.Y = self.model(parameters=self.current_state.cpu(), as_representation=True)
        sub = self.model.target[self.model.window]
        D = sub.flatten("data")
        V = sub.flatten("variance") if self.model.target.has_variance else 1.0
        r2 = (D - self.Y.flatten("data")) ** 2 / V
        if self.model.target.has_mask:
            r2 = r2[torch.logical_not(sub.flatten("mask"))]
        loss = float(torch.sum(r2).item()) / self.ndf
        if self.verbose > 0:
            AP_config.ap_logger.info(f"Loss: {loss}")
        self.lambda_history.append(self.current_state.detach().cpu().numpy().copy())
        self.loss_history.append(loss)
        if self.iteration >= 2 and (-self.relative_tolerance * 1e-3) < (
                (self.loss_history[-2] - self.loss_history[-1]) / self.loss_history[-1]) < (self.relative_tolerance / 10):
            self._count_finish += 1
        else:
            self._count_finish = 0
        self.iteration += 1

    def fit(self):
        self.iteration = 0
        self.Y = self.model(parameters=self.current_state.cpu(), as_representation=True)
        try:
            while True:
                self.step()
                if self.iteration > 2 and self._count_finish >= 2:
                    self.message = self.message + "success"
                    break
                elif self.iteration >= self.max_iter:
                    self.message = self.message + f"fail max iterations reached: {self.iteration}"
                    break
        except KeyboardInterrupt:
            self.message = self.message + "fail interrupted"
        self.model.parameters.vector_set_representation(self.res())
        return self


class Iter_LM(BaseOptimizer):
    """Levenberg-Marquardt on one chunk of the parameters at a time (drop-in for ``ap.fit.Iter_LM``, reference
    `fit/iterative.py:183-338`): the parameter vector is cut into chunks (``chunks``: a size, or explicit tuples of
    parameter identities), every chunk is fitted with the device LM while the others are held fixed
    (``Param_Mask``), and the sweep over all chunks repeats until chi^2 stops moving.  A chunk fit is one plan
    whose free parameters are the chunk: the fused normal equations only carry those columns."""

    def __init__(self, model, initial_state=None, chunks=50, max_iter=100, method="random", LM_kwargs=None, **kwargs):
        super().__init__(model, initial_state, max_iter=max_iter, **kwargs)
        self.max_iter = max_iter
        self.chunks = chunks
        self.method = method
        self.LM_kwargs = dict(LM_kwargs or {})
        sub = self.model.target[self.model.window]
        self.ndf = sub.flatten("data").numel() - len(self.current_state)
        if self.model.target.has_mask:
            self.ndf -= int(torch.sum(sub.flatten("mask")).item())
        self._count_finish = 0

    def _sweep(self, identities):
        """Boolean masks over ``identities``, one per chunk of this sweep, in the order the reference visits them
        (`fit/iterative.py:225-275`): an integer ``chunks`` deals the identities out ``chunks`` at a time (front of the
        remaining list, or ``random.sample`` of it), explicit chunks are taken in order or drawn without replacement."""
        import random

        n = len(identities)
        where = {pid: k for k, pid in enumerate(identities)}

        def mask_of(pids):
            m = torch.zeros(n, dtype=torch.bool)
            m[[where[p] for p in pids]] = True
            return m

        if isinstance(self.chunks, int):
            left = list(identities)
            while left:
                take = random.sample(left, min(len(left), self.chunks)) if self.method == "random" else left[: self.chunks]
                m = mask_of(take)
                left = [pid for pid in left if not m[where[pid]]]
                yield m
        elif isinstance(self.chunks, (tuple, list)):
            left = list(range(len(self.chunks)))
            while left:
                k = random.choice(left) if self.method == "random" else left[0]
                left.remove(k)
                yield mask_of(self.chunks[k])
        else:
            raise ValueError(f"Unrecognized chunks value, should be one of int, tuple. not: {type(self.chunks)}")

    def step(self):
        """One sweep: every chunk fitted once with the others held fixed, then the bookkeeping of the sweep."""
        from .param import Param_Mask

        res = None
        if self.verbose > 0:
            AP_config.ap_logger.info("--------iter-------")
        for chunk in self._sweep(list(self.model.parameters.vector_identities())):
            if self.verbose > 1:
                AP_config.ap_logger.info(str(chunk))
            with Param_Mask(self.model.parameters, chunk):
                res = LM(self.model, ndf=self.ndf, **self.LM_kwargs).fit()
            if self.verbose > 0:
                AP_config.ap_logger.info(f"chunk loss: {res.res_loss()}")
        self.loss_history.append(res.res_loss())
        self.lambda_history.append(self.model.parameters.vector_representation().detach().cpu().numpy())
        gain = (self.loss_history[-2] - self.loss_history[-1]) / self.loss_history[-1] if self.iteration >= 2 else None
        flat = gain is not None and -self.relative_tolerance * 1e-3 < gain < self.relative_tolerance / 10
        self._count_finish = self._count_finish + 1 if flat else 0
        self.iteration += 1

    def fit(self):
        self.iteration = 0
        try:
            while True:
                self.step()
                if self.iteration > 2 and self._count_finish >= 2:
                    self.message = self.message + "success"
                    break
                elif self.iteration >= self.max_iter:
                    self.message = self.message + f"fail max iterations reached: {self.iteration}"
                    break
        except KeyboardInterrupt:
            self.message = self.message + "fail interrupted"
        self.model.parameters.vector_set_representation(self.res())
        return self
