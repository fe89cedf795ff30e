"""How even are the tiles of the crowded field?  CPU study behind lowering._tile_cost (no GPU needed).

For every Sersic source of config[2] the oracle's refinement is replayed with a per-pixel count of profile evaluations
(first pass + every Gauss-Legendre node of every queue entry, all depths): the work k_first / k_integrate_pool do.  A
tile pays for the pixels of a source's evaluation region (window + PSF border) that fall on (tile + PSF border).  Prints
max / mean tile load for even cuts and for the cuts tile_scene makes.
    python scripts/tile_balance_study.py [c3|c3s]
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import astrophot_b200 as ap
import astrophot_oracle as orc
from astrophot_b200 import scene as sc
from astrophot_b200.lowering import lower, tile_scene
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
ap.AP_config.ap_device = "cpu"
model = bench.build_workload(ap, wl, 1, None)
scene, _ = lower(model, for_fit=True)
im = scene.images[0]
vals = orc.rep_to_val(model.parameters.vector_representation().numpy(), scene.transform, scene.lo, scene.hi)[0]
S = np.asarray(im.S, dtype=np.float64)
area = abs(np.linalg.det(S))
cache = os.path.join("/tmp/study", f"cost_{wl}.npz")


def count_map(si):
    """Evaluations per pixel of the evaluation region of source si (value pass)."""
    src = scene.sources[si]
    el = orc.source_elements(src, vals)
    psf = orc._host(scene.psfs[src.psf].data)
    bx, by = orc._psf_border(psf)
    ox, oy, ow, oh = src.out
    ex0, ey0, ew, eh = ox - bx, oy - by, ow + 2 * bx, oh + 2 * by
    Sinv = np.linalg.inv(S)
    pc = Sinv @ (np.array([el[0], el[1]]) - im.rxy) + im.rij
    rnd = np.round(pc)
    jj, ii = np.meshgrid(np.arange(ey0 - 1, ey0 + eh + 1, dtype=np.float64), np.arange(ex0 - 1, ex0 + ew + 1, dtype=np.float64), indexing="ij")
    X, Y = S[0, 0] * (ii - rnd[0]) + S[0, 1] * (jj - rnd[1]), S[1, 0] * (ii - rnd[0]) + S[1, 1] * (jj - rnd[1])
    deep, _ = orc.eval_profile(src, el, X, Y, area, False)
    lap = np.abs(deep[:-2, 1:-1] + deep[2:, 1:-1] + deep[1:-1, :-2] + deep[1:-1, 2:] - 4 * deep[1:-1, 1:-1])
    rw, rh = im.W + 2 * bx, im.H + 2 * by        # inside a group the working window is the group's
    thr = src.tolerance * orc.sersic_total_flux(10.0 ** el[6], el[4], el[5], el[2]) / (rw * rh)
    cnt = np.ones((eh, ew))
    sel = lap > thr
    Xe, Ye = X[1:-1, 1:-1], Y[1:-1, 1:-1]
    pix = np.flatnonzero(sel.reshape(-1))
    x, y, s_, a_, t = Xe.reshape(-1)[pix], Ye.reshape(-1)[pix], S, area, thr
    c = cnt.reshape(-1)
    for depth in range(1, src.max_depth + 1):
        if x.size == 0:
            break
        res, ref, _ = orc._gl(src, el, x, y, s_, a_, src.quad_level, False)
        np.add.at(c, pix, src.quad_level ** 2)
        if depth == src.max_depth:
            break
        more = np.abs(res - ref) > t
        N = src.gridding
        sx, sy = orc.sub_offsets(N, s_)
        x = (x[more][:, None] + sx).reshape(-1)
        y = (y[more][:, None] + sy).reshape(-1)
        pix = np.repeat(pix[more], N * N)
        s_, a_, t = s_ / N, a_ / N ** 2, t * N ** 2
    return (ex0, ey0, cnt)


gal = [si for si, s in enumerate(scene.sources) if s.kind == sc.KIND_SERSIC]
if os.path.exists(cache):
    z = np.load(cache, allow_pickle=True)
    maps = list(z["maps"])
else:
    orc.set_threads(8)
    maps = orc._pmap(count_map, gal)
    np.savez(cache, maps=np.array(maps, dtype=object))
tot = sum(m[2].sum() for m in maps)
print(f"{wl}: {len(gal)} Sersic sources, {tot:.3e} evaluations per value pass")
psf = orc._host(scene.psfs[0].data)
bx, by = orc._psf_border(psf)


def loads(tiled):
    out = []
    for t in tiled.images:
        x0, y0 = (np.round(np.asarray(im.rij) - np.asarray(t.rij))).astype(int)
        out.append([x0, y0, t.W, t.H, 0.0])
    for (ex0, ey0, cnt), si in zip(maps, gal):
        ox, oy, ow, oh = scene.sources[si].out
        for L in out:
            x0, y0, w, h, _ = L
            # the piece of the source on this tile: out window clipped to the tile, evaluated with the PSF border
            cx0, cy0, cx1, cy1 = max(ox, x0), max(oy, y0), min(ox + ow, x0 + w), min(oy + oh, y0 + h)
            if cx1 <= cx0 or cy1 <= cy0:
                continue
            a0, b0, a1, b1 = cx0 - bx - ex0, cy0 - by - ey0, cx1 + bx - ex0, cy1 + by - ey0
            L[4] += cnt[b0:b1, a0:a1].sum()
    return np.array([L[4] for L in out])


for ny, nx in ((1, 2), (2, 2), (2, 4)):
    for bal in (False, True):
        l = loads(tile_scene(scene, ny, nx, balance=bal))
        print(f"  {ny}x{nx} balance={bal}: max/mean = {l.max() / l.mean():.3f}, sum/total = {l.sum() / tot:.3f}, "
              f"max/(total/N) = {l.max() / (tot / (ny * nx)):.3f}")
