"""CPU: host-side behaviour of the training-step modules (no kernels run): argument validation, the
no-CPU-fallback rule, optimizer state layout, learning-rate schedule."""
import math

import pytest
import torch


def test_image_loss_rejects_cpu_tensors(built_lib):
    from s3gaussian_b200 import losses
    with pytest.raises(RuntimeError, match="no CPU path"):
        losses.l1_loss(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8))
    with pytest.raises(RuntimeError, match="pairs"):
        losses.image_loss_terms(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8), depth=torch.zeros(1, 8, 8))
    with pytest.raises(NotImplementedError):
        losses.compute_depth("smooth_l1", torch.zeros(1, 8, 8), torch.zeros(1, 8, 8))
    with pytest.raises(NotImplementedError):
        losses.ssim(torch.zeros(3, 8, 8), torch.zeros(3, 8, 8), window_size=7)


def test_fused_adam_rejects_cpu_and_unsupported(built_lib):
    from s3gaussian_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FusedAdam([p], lr=1e-3).step()
    with pytest.raises(NotImplementedError):
        FusedAdam([p], lr=1e-3, amsgrad=True)
    with pytest.raises(NotImplementedError):
        FusedAdam([p], lr=1e-3, weight_decay=0.1)
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def test_fused_adam_keeps_the_reference_group_layout(built_lib):
    """The reference addresses groups by group['name'] and state by parameter (gaussian_model.py:397-470)."""
    from s3gaussian_b200.optim import FusedAdam
    a, b = torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(5, 1))
    opt = FusedAdam([{"params": [a], "lr": 1.6e-4, "name": "xyz"}, {"params": [b], "lr": 0.05, "name": "opacity"}],
                    lr=0.0, eps=1e-15)
    assert [g["name"] for g in opt.param_groups] == ["xyz", "opacity"]
    assert opt.param_groups[0]["eps"] == 1e-15 and opt.param_groups[1]["lr"] == 0.05
    assert opt.state.get(a, None) is None
    opt.step()      # no gradients: nothing to do, no state created, no kernel launched
    assert len(opt.state) == 0
    sd = opt.state_dict()
    assert sd["param_groups"][0]["name"] == "xyz"


def test_densify_stats_rejects_cpu(built_lib):
    from s3gaussian_b200.optim import add_densification_stats
    with pytest.raises(RuntimeError, match="no CPU path"):
        add_densification_stats(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32), torch.zeros(4, 1),
                                torch.zeros(4, 1), torch.zeros(4))


def test_expon_lr_schedule():
    from s3gaussian_b200.optim import get_expon_lr_func
    f = get_expon_lr_func(1.6e-4, 1.6e-6, lr_delay_mult=0.01, max_steps=20000)
    assert f(0) == pytest.approx(1.6e-4) and f(20000) == pytest.approx(1.6e-6) and f(40000) == pytest.approx(1.6e-6)
    assert f(10000) == pytest.approx(math.sqrt(1.6e-4 * 1.6e-6))
    assert f(-1) == 0.0 and get_expon_lr_func(0.0, 0.0)(5) == 0.0
    d = get_expon_lr_func(1.0, 1.0, lr_delay_steps=100, lr_delay_mult=0.1)
    assert d(0) == pytest.approx(0.1) and d(100) == pytest.approx(1.0) and 0.1 < d(50) < 1.0


def test_split_plan_row_order():
    """kept rows in order, then the selected rows twice - the order densify_and_split + prune_points produce
    (scene/gaussian_model.py:496-522)."""
    from s3gaussian_b200.gaussian_model import split_plan
    sel = torch.tensor([False, True, False, False, True, False])
    src, n_kept, idx = split_plan(sel)
    assert n_kept == 4 and idx.tolist() == [1, 4]
    assert src.tolist() == [0, 2, 3, 5, 1, 4, 1, 4]
    # the same thing spelled the reference's way on a value tensor
    x = torch.arange(6.0)
    cat = torch.cat((x, x[sel].repeat(2)))
    pruned = cat[~torch.cat((sel, torch.zeros(2 * int(sel.sum()), dtype=torch.bool)))]
    assert torch.equal(pruned, x[src])


def test_gaussian_model_rejects_cpu_and_exposes_reference_names(built_lib):
    from s3gaussian_b200.gaussian_model import GaussianModel, default_optimization_params
    m = GaussianModel(3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.create_from_tensors(torch.zeros(4, 3), torch.zeros(4, 1, 3), torch.zeros(4, 15, 3), torch.zeros(4, 3),
                              torch.zeros(4, 4), torch.zeros(4, 1))
    for name in ("create_from_pcd", "save_ply", "load_ply", "save_deformation", "load_model","training_setup", "update_learning_rate", "add_densification_stats", "densify", "prune", "prune_points",
                 "densify_and_clone", "densify_and_split", "reset_opacity", "replace_tensor_to_optimizer",
                 "compute_regulation", "oneupSHdegree", "get_covariance"):
        assert callable(getattr(m, name))
    a = default_optimization_params()
    assert a.position_lr_init == 0.00016 and a.percent_dense == 0.01 and a.opacity_lr == 0.05


def test_gather_rows_argument_checks(built_lib):
    import ctypes as C
    from s3gaussian_b200 import _lib
    lib = _lib.load()
    assert lib.s3g_gather_rows(0, None, 1, 1, None, None) == -1
    t = (_lib.RowTensor * 1)(_lib.RowTensor(1, 1, 3, 0))
    assert lib.s3g_gather_rows(1, t, 4, 5, None, None) == -1          # n_kept > n_out
    assert lib.s3g_gather_rows(1, t, 0, 0, None, None) == 0           # nothing to do


def test_peer_slices_cover_the_buffer_in_float4_units():
    from s3gaussian_b200.dp import peer_slices
    for numel, world in ((40, 3), (8, 8), (118_000_000, 8), (4, 2), (472, 5)):
        sl = peer_slices(numel, world)
        assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == numel
        assert all(b % 4 == 0 and e % 4 == 0 and b <= e for b, e in sl)
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        peer_slices(10, 2)


def test_peer_entry_points_check_arguments(built_lib):
    import ctypes as C
    from s3gaussian_b200 import _lib
    lib = _lib.load()
    two = (C.c_void_p * 2)(16, 32)
    assert lib.s3g_peer_reduce_scatter(1, 0, two, 8, None) == -1        # world < 2
    assert lib.s3g_peer_reduce_scatter(2, 2, two, 8, None) == -1        # rank out of range
    assert lib.s3g_peer_all_gather(2, 0, two, 6, None) == -1            # numel not a multiple of 4
    bad = (C.c_void_p * 2)(16, 36)
    assert lib.s3g_peer_all_gather(2, 0, bad, 8, None) == -1            # misaligned peer pointer
