"""GPU: the optional FCNet layouts of the ModelV2 class (`vf_share_layers`, `free_log_std`) on the two-branch kernels."""
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_modelv2 import MODEL_CONFIG, TOL, _O
from tests.util import scaled_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vf_share,free_std", [(True, False), (False, True), (True, True)])
def test_fcnet_modelv2_optional_layouts(vf_share, free_std):
    """`vf_share_layers` (value_out reads the policy branch, models/fcnet_glorot_uniform_init.py:95-113) and `free_log_std`
    (state-independent log-std variable, :30-36,85-93): variables in the reference's order, forward / value / gradients
    against the oracle's restatement of those layouts."""
    from ddrl_b200 import spaces
    from ddrl_b200.catalog import ModelCatalog
    import ddrl_b200.modelv2  # noqa: F401
    O = _O()
    D, A = 27, 4
    cfg = dict(MODEL_CONFIG, custom_model="ffn", vf_share_layers=vf_share, free_log_std=free_std)
    model = ModelCatalog.get_model_v2(spaces.Box(-np.inf, np.inf, (D,), np.float64), spaces.Box(-1.0, 1.0, (A,)), 2 * A, cfg,
                                      name="p")
    shapes = O.fcnet_shapes(D, 2 * A, (64, 64), vf_share, free_std)
    assert list(model.variables(as_dict=True)) == ["p/" + n for n, _ in shapes]
    assert [tuple(v.shape) for v in model.variables()] == [tuple(sh) for _, sh in shapes]
    assert model.theta.numel() == O.n_params(shapes)
    g = torch.Generator().manual_seed(5)
    theta = O.fcnet_init(D, 2 * A, g, vf_share_layers=vf_share, free_log_std=free_std, dtype=torch.float64)
    theta = (theta + 0.05 * torch.randn(theta.shape, generator=g, dtype=torch.float64)).float()
    model.load_flat(theta)
    x = torch.randn(300, D, generator=g)
    out, _ = model({"obs_flat": x.cuda()}, [], None)
    val = model.value_function()
    t = theta.double().requires_grad_(True)
    lg, vr = O.fcnet_forward(t, x.double(), 2 * A, vf_share_layers=vf_share, free_log_std=free_std)
    assert scaled_err(out.detach().cpu().numpy(), lg.detach().numpy()) < TOL
    assert scaled_err(val.detach().cpu().numpy(), vr.detach().numpy()) < TOL
    if free_std:      # the second half of the logits is the log_std variable, identical for every row
        assert torch.equal(out[:, A:].detach().cpu(), theta[:A].expand(300, A))
    wl, wv = torch.randn(300, 2 * A, generator=g), torch.randn(300, generator=g)
    (gd,) = torch.autograd.grad((out * wl.cuda()).sum() + (val * wv.cuda()).sum(), model.theta)
    (ref,) = torch.autograd.grad((lg * wl.double()).sum() + (vr * wv.double()).sum(), t)
    assert gd.shape == model.theta.shape
    assert scaled_err(gd.cpu().numpy(), ref.numpy()) < 2 * TOL       # tied weights: two FP32 branch gradients are added
    o = 0
    for name, shp in shapes:      # every variable at its own scale: a wiring error (tying, constant columns) would be O(1)
        n = int(np.prod(shp))
        assert scaled_err(gd[o:o + n].cpu().numpy(), ref[o:o + n].numpy()) < 1e-4, name
        o += n
