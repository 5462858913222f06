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
