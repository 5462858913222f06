"""CPU: size-independent properties of the rasterizer oracle on seeded random scenes - the same properties the
GPU suite checks on the CUDA path at full size (tests/test_gpu_raster.py::test_full_size_properties), here as a
guard on the checker itself: key order, range partition, contributor bounds, linearity of the backward in the
incoming image gradient, empty and fully culled inputs."""
import numpy as np
import pytest
import torch

import util


def _scene(P, W, H, seed, mode="sh"):
    from s3gaussian_b200 import synthetic as syn
    cloud, cam = syn.make_small_scene(P=P, width=W, height=H, seed=seed)
    return util.scene_inputs(cloud, cam, mode=mode, sh_degree=3)


@pytest.mark.parametrize("P,W,H,seed,mode", [(500, 96, 64, 1, "sh"), (900, 70, 50, 2, "rgb"), (64, 33, 17, 3, "sh")])
def test_binning_invariants(P, W, H, seed, mode, oracle_lib):
    d = _scene(P, W, H, seed, mode)
    o = util.oracle_run(oracle_lib, d)
    keys = o["binning"]["keys"].astype(np.uint64)
    pl = o["binning"]["point_list"]
    R = o["R"]
    assert len(keys) == R == int(o["geometry"]["tiles_touched"][o["radii"] > 0].sum())
    assert np.all(keys[1:] >= keys[:-1])                                   # sorted by (tile << 32 | depth bits)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    ranges = o["binning"]["ranges"].reshape(-1, 2).astype(np.int64)
    tx, ty = (W + 15) // 16, (H + 15) // 16
    assert ranges.shape[0] == tx * ty
    covered = 0
    for t, (b, e) in enumerate(ranges):
        if e > b:
            assert np.all(tiles[b:e] == t)
            covered += e - b
    assert covered == R                                                    # the ranges partition the list
    assert np.all(o["radii"][pl] > 0)                                      # only visible Gaussians are listed
    # every pixel's contributor count is bounded by its tile's list length
    n_contrib = o["image"]["n_contrib"].reshape(H, W)
    lens = (ranges[:, 1] - ranges[:, 0]).reshape(ty, tx)
    per_pixel = np.repeat(np.repeat(lens, 16, axis=0), 16, axis=1)[:H, :W]
    assert np.all(n_contrib <= per_pixel)
    T = o["image"]["final_T"].reshape(H, W)
    assert np.all((T >= 0) & (T <= 1.0 + 1e-6))


def test_backward_is_linear_in_the_image_gradient(oracle_lib):
    d = _scene(400, 64, 48, 7)
    gc, gd = util.seeded_grads(d, 3)
    g1 = util.oracle_run(oracle_lib, d, gc, gd)["grads"]
    g2 = util.oracle_run(oracle_lib, d, 2.0 * gc, 2.0 * gd)["grads"]
    hc, hd = util.seeded_grads(d, 4)
    g3 = util.oracle_run(oracle_lib, d, hc, hd)["grads"]
    g13 = util.oracle_run(oracle_lib, d, gc + hc, gd + hd)["grads"]
    for k in g1:
        if g1[k] is None:
            continue
        a1, a2, a3, a13 = (np.asarray(x[k], np.float64) for x in (g1, g2, g3, g13))
        scale = np.abs(a1).max() + np.abs(a3).max() + 1e-30
        assert np.abs(a2 - 2 * a1).max() <= 1e-5 * scale, k
        assert np.abs(a13 - (a1 + a3)).max() <= 1e-4 * scale, k


def test_empty_and_fully_culled(oracle_lib):
    d = _scene(50, 48, 32, 5)
    behind = dict(d)
    m = d["means3D"].clone()
    m[:, 2] = -abs(m[:, 2]) - 5.0                                          # behind the synthetic camera's near plane
    behind["means3D"] = m
    o = util.oracle_run(oracle_lib, behind)
    if o["R"] == 0:                                                        # camera looks down +z in make_small_scene
        assert np.all(o["radii"] == 0)
        assert np.allclose(o["color"], d["bg"].numpy()[:, None, None])
        assert np.all(o["image"]["n_contrib"] == 0)
    # whatever the camera convention, a cloud with zero opacity-weight leaves the background untouched
    dim = dict(d)
    dim["opacities"] = torch.zeros_like(d["opacities"])
    o2 = util.oracle_run(oracle_lib, dim)
    assert np.allclose(o2["color"], d["bg"].numpy()[:, None, None], atol=1e-7)
    assert np.all(o2["image"]["n_contrib"] == 0)
