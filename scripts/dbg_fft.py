import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import astrophot_b200 as ap, scenes
import astrophot_oracle as orc
from astrophot_b200.lowering import lower
from astrophot_b200.cabi import Plan
from conftest import load_golden
fix = load_golden("crowded")
model, _ = scenes.build(ap, "crowded")
scene, info = lower(model)
w = orc.sample(scene, fix["x_val"], as_rep=False)[0]
for conv in ("direct", "fft", "direct", "fft"):
    pl = Plan(scene, conv=conv)
    a = pl.sample(fix["x_val"])[0].cpu().numpy()
    print(conv, "vs oracle", np.abs(a - w).max(), "stats", pl.stats())
    a2 = pl.sample(fix["x_val"])[0].cpu().numpy()
    print(conv, "2nd call vs oracle", np.abs(a2 - w).max())
    del pl
def sub(names):
    return type(scene)(images=scene.images, sources=[s for s in scene.sources if s.name in names], psfs=scene.psfs,
                       transform=scene.transform, lo=scene.lo, hi=scene.hi, identities=scene.identities)
gals = [f"g{k}" for k in range(8)]
pts = [f"p{k}" for k in range(14)]
for names in (gals, gals + ["csky"], gals + pts[:1], gals[:1] + pts[:1], pts, gals[:2] + pts):
    sc2 = sub(names)
    ww = orc.sample(sc2, fix["x_val"], as_rep=False)[0]
    for conv in ("direct", "fft"):
        a = Plan(sc2, conv=conv).sample(fix["x_val"])[0].cpu().numpy()
        print(len(names), names[-1], conv, "vs oracle", np.abs(a - ww).max())
