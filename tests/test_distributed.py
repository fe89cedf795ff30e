"""N>1 path on CPU: two gloo ranks each take their share of the bands of a joint
fit (lowering.shard_scene), build local normal equations (CPU oracle standing in
for the kernels), all-reduce them, and must reproduce the unsharded result — the
same reduction LM(distributed=True) performs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, golden_data


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    import scenes
    from astrophot_b200.lowering import lower, shard_scene

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ap.AP_config.ap_device = "cpu"
    fix = load_golden("joint")
    model, _ = scenes.build(ap, "joint", data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    local = shard_scene(scene, rank, world)
    assert len(local.images) == len([i for i in range(3) if i % world == rank])
    assert local.n_par == scene.n_par
    H, g, chi2, _ = orc.normal_eq(local, fix["x0"])
    buf = torch.cat([torch.as_tensor(H).reshape(-1), torch.as_tensor(g), torch.tensor([chi2])])
    dist.all_reduce(buf)
    n_keep = torch.tensor([float(sum(im.H * im.W for im in local.images))], dtype=torch.float64)
    dist.all_reduce(n_keep)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), np.concatenate([buf.numpy(), n_keep.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_band_sharded_normal_equations_allreduce(tmp_path):
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    import scenes
    from astrophot_b200.lowering import lower

    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    red = np.load(tmp_path / "reduced.npy")
    fix = load_golden("joint")
    P = len(fix["x0"])
    H, g, chi2, npix = red[: P * P].reshape(P, P), red[P * P : P * P + P], red[P * P + P], red[-1]
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H - fix["hess0"]) / np.outer(d, d)) < 1e-9      # reference's own J^T W J
    assert np.max(np.abs(g - fix["grad0"])) / np.abs(fix["grad0"]).max() < 1e-9
    assert npix == 3 * 48 * 48
    assert abs(chi2 / (npix - P) - fix["loss_history"][0]) / fix["loss_history"][0] < 1e-10


# ---------------------------------------------------------------------------
# tile sharding of ONE image: lowering.tile_scene
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,tiles", [("crowded", (2, 2)), ("crowded", (2, 3)), ("group", (1, 2)), ("group_nosky", (2, 1))])
def test_tiles_reproduce_the_whole_image(name, tiles):
    """Pixels owned once, sources handed to every tile they touch: model image, J^T W J, J^T W r and
    chi^2 of the tiles stitch / add up to those of the whole image and to the reference's golden."""
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    import scenes
    from astrophot_b200.lowering import lower, tile_scene

    ap.AP_config.ap_device = "cpu"
    fix = load_golden(name)
    model, _ = scenes.build(ap, name, data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    tiled = tile_scene(scene, *tiles)
    assert len(tiled.images) == tiles[0] * tiles[1] * len(scene.images)
    assert sum(im.H * im.W for im in tiled.images) == sum(im.H * im.W for im in scene.images)
    x0 = fix["x0"]
    whole = orc.sample(scene, x0)[0]
    parts = orc.sample(tiled, x0)
    # a tile's origin on the whole image is the shift of its reference pixel (the cuts are cost-balanced, not even)
    stitched = np.full_like(whole, np.nan)
    for im, part in zip(tiled.images, parts):
        x0t, y0t = (np.round(np.asarray(scene.images[0].rij) - np.asarray(im.rij))).astype(int)
        assert np.all(np.isnan(stitched[y0t:y0t + im.H, x0t:x0t + im.W]))      # every pixel owned once
        stitched[y0t:y0t + im.H, x0t:x0t + im.W] = part
    assert not np.any(np.isnan(stitched))
    even = tile_scene(scene, *tiles, balance=False)
    assert [(im.H, im.W) for im in even.images] == [
        (round((a + 1) * scene.images[0].H / tiles[0]) - round(a * scene.images[0].H / tiles[0]),
         round((b + 1) * scene.images[0].W / tiles[1]) - round(b * scene.images[0].W / tiles[1]))
        for a in range(tiles[0]) for b in range(tiles[1])]
    assert np.max(np.abs(stitched - whole)) <= 1e-13 * np.max(np.abs(whole))
    H, g, chi2, _ = orc.normal_eq(tiled, x0)
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(g - fix["grad0"])) / np.abs(fix["grad0"]).max() < 1e-9


def _tile_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import astrophot_b200 as ap
    import astrophot_oracle as orc
    import scenes
    from astrophot_b200.lowering import lower, shard_scene, tile_scene

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ap.AP_config.ap_device = "cpu"
    fix = load_golden("crowded")
    model, _ = scenes.build(ap, "crowded", data=golden_data(fix))
    scene, _ = lower(model, for_fit=True)
    local = shard_scene(tile_scene(scene, 2, 2), rank, world)
    assert len(local.images) == 2 and local.n_par == scene.n_par
    H, g, chi2, _ = orc.normal_eq(local, fix["x0"])
    buf = torch.cat([torch.as_tensor(H).reshape(-1), torch.as_tensor(g), torch.tensor([chi2])])
    dist.all_reduce(buf)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced_tiles.npy"), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_tile_sharded_normal_equations_allreduce(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_tile_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    red = np.load(tmp_path / "reduced_tiles.npy")
    fix = load_golden("crowded")
    P = len(fix["x0"])
    H, g, chi2 = red[: P * P].reshape(P, P), red[P * P : P * P + P], red[P * P + P]
    d = np.sqrt(np.diag(fix["hess0"]))
    assert np.max(np.abs(H - fix["hess0"]) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(g - fix["grad0"])) / np.abs(fix["grad0"]).max() < 1e-9
    npix = 192 * 192
    assert abs(chi2 / (npix - P) - fix["loss_history"][0]) / fix["loss_history"][0] < 1e-10


# ---------------------------------------------------------------------------
# the whole distributed LM (fit.LM(distributed=True, tiles=...)) on CPU: gloo ranks, oracle-backed stand-in plans
# ---------------------------------------------------------------------------
def _lm_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import astrophot_b200 as ap
    import scenes
    from astrophot_b200 import cabi
    from test_lm_host_logic import OraclePlan, _solve

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ap.AP_config.ap_device = "cpu"
    cabi.Plan, cabi.lm_solve = OraclePlan, _solve
    fix = load_golden("group")
    model, _ = scenes.build(ap, "group", data=golden_data(fix))
    res = ap.fit.LM(model, initial_state=fix["x0"], max_iter=4, relative_tolerance=0.0, distributed=True, tiles=(2, 1),
                    fused_trial=False).fit()
    assert len(res.plan.shapes) == 1              # one of the two tiles
    if rank == 0:
        np.savez(os.path.join(out_dir, "lm_tiles.npz"), loss=np.array(res.loss_history), L=np.array(res.L_history),
                 lam=np.array(res.lambda_history))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_tiled_lm_control_flow(tmp_path):
    """Two gloo ranks, one image tile each: the ranks' normal equations, geodesic right-hand sides and chi^2 records
    are all-reduced where fit.LM says, every rank takes the same decisions, and the history is the reference's."""
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_lm_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "lm_tiles.npz")
    fix = load_golden("group")
    n = len(got["loss"])
    np.testing.assert_allclose(got["loss"], fix["loss_history"][:n], rtol=1e-8)
    np.testing.assert_allclose(got["L"][:4], fix["L_history"][:4], rtol=1e-12)
    np.testing.assert_allclose(got["lam"], fix["lambda_history"][:n], rtol=1e-8, atol=1e-8)


def test_more_ranks_than_images_is_refused_on_every_rank():
    """shard_scene decides from the whole scene: with more ranks than bands / tiles every rank raises the same error
    (a rank left without images would otherwise stop alone and the others would hang in their first collective)."""
    import astrophot_b200 as ap
    import scenes
    from astrophot_b200.lowering import lower, shard_scene
    ap.AP_config.ap_device = "cpu"
    model, _ = scenes.build(ap, "joint")
    scene, _ = lower(model)
    n_img = len([im for im in scene.images if not im.aux])
    for rank in range(n_img + 1):
        with pytest.raises(ap.errors.SpecificationConflict, match="at least"):
            shard_scene(scene, rank, n_img + 1)
    assert len(shard_scene(scene, 0, n_img).images) == 1
