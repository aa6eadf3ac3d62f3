"""GPU: the RLlib ModelV2 mirror classes (same names / constructor / forward / value_function as models/*.py) run
through the C-ABI kernels, match the oracle, and carry gradients through torch.autograd to their flat parameters."""
import numpy as np
import pytest
import torch

from tests.util import ckpt_theta, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
MODEL_CONFIG = {"fcnet_hiddens": [64, 64], "fcnet_activation": "tanh", "free_log_std": False, "no_final_linear": False,
                "vf_share_layers": False}


def _O():
    import oracle.ddrl_oracle as O
    return O


def test_fcnet_modelv2_forward_value_and_autograd():
    from ddrl_b200 import spaces
    from ddrl_b200.catalog import ModelCatalog
    import ddrl_b200.modelv2  # noqa: F401  (registers the models)
    O = _O()
    theta, _, D, A = ckpt_theta("FullyDecentral")
    cfg = dict(MODEL_CONFIG, custom_model="fc_glorot_uniform_init")
    model = ModelCatalog.get_model_v2(spaces.Box(-np.inf, np.inf, (D,), np.float64), spaces.Box(-1.0, 1.0, (A,)), 2 * A,
                                      cfg, name="policy_FL")
    assert type(model).__name__ == "FullyConnectedNetwork_GlorotUniformInitializer"
    # variable names / order of the reference checkpoint
    names = list(model.variables(as_dict=True))
    assert names[:4] == ["policy_FL/fc_1/kernel", "policy_FL/fc_1/bias", "policy_FL/fc_value_1/kernel", "policy_FL/fc_value_1/bias"]
    assert names[-2:] == ["policy_FL/value_out/kernel", "policy_FL/value_out/bias"]
    # Glorot init: output layers are 10x smaller (scale 0.01), biases zero
    v = model.variables(as_dict=True)
    assert float(v["policy_FL/fc_1/bias"].abs().max()) == 0.0
    assert float(v["policy_FL/fc_out/kernel"].abs().max()) <= np.sqrt(0.06 / 68.0) + 1e-7
    assert float(v["policy_FL/fc_1/kernel"].abs().max()) <= np.sqrt(6.0 / 83.0) + 1e-7
    model.load_flat(theta[0])                                    # checkpoint import = one copy
    x = torch.randn(200, D)
    out, state = model({"obs": x.cuda(), "obs_flat": x.cuda()}, [], None)
    val = model.value_function()
    lg, vr = O.fcnet_forward(torch.from_numpy(theta[0]).double(), x.double(), 2 * A)
    assert state == [] and out.shape == (200, 2 * A) and val.shape == (200,)
    assert scaled_err(out.detach().cpu().numpy(), lg.numpy()) < TOL
    assert scaled_err(val.detach().cpu().numpy(), vr.numpy()) < TOL
    # autograd to the flat parameter vector
    wl, wv = torch.randn(200, 2 * A), torch.randn(200)
    loss = (out * wl.cuda()).sum() + (val * wv.cuda()).sum()
    (g,) = torch.autograd.grad(loss, model.theta)
    t = torch.from_numpy(theta[0]).double().requires_grad_(True)
    lg, vr = O.fcnet_forward(t, x.double(), 2 * A)
    (ref,) = torch.autograd.grad((lg * wl.double()).sum() + (vr * wv.double()).sum(), t)
    assert scaled_err(g.cpu().numpy(), ref.numpy()) < TOL
    assert ModelCatalog.get_custom_model("ffn") is type(model)


def test_gnn_modelv2_forward_and_autograd():
    from ddrl_b200 import spaces
    from ddrl_b200.catalog import ModelCatalog
    from ddrl_b200.policies import QuantrupedDecentralizedSharedGraphEnv as Env
    import ddrl_b200.modelv2  # noqa: F401
    O = _O()
    (_, obs_space, act_space, _), = Env.return_policies().values()
    model = ModelCatalog.get_model_v2(obs_space, act_space, 4, dict(MODEL_CONFIG, custom_model="gnn"), name="leg_policy")
    assert type(model).__name__ == "FullyConnectedNetwork_GNN_GlorotUniformInitializer"
    th = model.theta.detach().cpu()
    assert th.numel() == 28869
    B = 150
    rng = np.random.default_rng(0)
    idx = torch.from_numpy(rng.integers(0, 4, size=(B, 1)))
    state = torch.randn(B, 4, 23)
    adj = torch.from_numpy(np.broadcast_to(Env.create_adj(), (B, 4, 4)).copy()).float()
    out, _ = model({"obs": (idx.cuda(), state.cuda(), adj.cuda())}, [], None)
    val = model.value_function()
    lg, vr = O.graphnet_forward(th.double(), idx, state.double(), adj.double(), 4)
    assert scaled_err(out.detach().cpu().numpy(), lg.numpy()) < TOL
    assert scaled_err(val.detach().cpu().numpy(), vr.numpy()) < TOL
    wl, wv = torch.randn(B, 4), torch.randn(B)
    (g,) = torch.autograd.grad((out * wl.cuda()).sum() + (val * wv.cuda()).sum(), model.theta)
    t = th.double().requires_grad_(True)
    lg, vr = O.graphnet_forward(t, idx, state.double(), adj.double(), 4)
    (ref,) = torch.autograd.grad((lg * wl.double()).sum() + (vr * wv.double()).sum(), t)
    assert scaled_err(g.cpu().numpy(), ref.numpy()) < 2e-5


def test_coupling_modelv2():
    from ddrl_b200 import spaces
    from ddrl_b200.catalog import ModelCatalog
    import ddrl_b200.modelv2  # noqa: F401
    O = _O()
    obs_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (19,), np.float64)])
    model = ModelCatalog.get_model_v2(obs_space, spaces.Box(-1.0, 1.0, (2,)), 4, dict(MODEL_CONFIG, custom_model="cup"),
                                      name="policy_legs")
    th = model.theta.detach().cpu()
    B = 64
    idx = torch.randint(0, 4, (B, 1))
    x = torch.randn(B, 19)
    with torch.no_grad():
        out, _ = model({"obs": (idx.cuda(), x.cuda())}, [], None)
    lg, vr = O.fcnet_forward(th.double(), x.double(), 4)
    ref = O.leg_coupling(lg, idx, torch.tensor(O.COUPLING_INIT, dtype=torch.float64))
    assert scaled_err(out.cpu().numpy(), ref.numpy()) < TOL
    assert scaled_err(model.value_function().cpu().numpy(), vr.numpy()) < TOL
    # with autograd the multiply stays on the graph and gives the same values
    out2, _ = model({"obs": (idx.cuda(), x.cuda())}, [], None)
    assert torch.allclose(out2.detach(), out, rtol=0, atol=1e-7)


def test_coupling_table_is_trainable():
    """The reference registers `leg_coupling` as a trainable tf.Variable (models/coupling_net_glorot_uniform_init.py:20-21,
    160-161): gradients w.r.t. the table (non-initial values) and w.r.t. the MLP weights against the oracle's autograd."""
    from ddrl_b200 import spaces
    from ddrl_b200.catalog import ModelCatalog
    import ddrl_b200.modelv2  # noqa: F401
    O = _O()
    obs_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (19,), np.float64)])
    model = ModelCatalog.get_model_v2(obs_space, spaces.Box(-1.0, 1.0, (2,)), 4, dict(MODEL_CONFIG, custom_model="cup"),
                                      name="policy_legs")
    g = torch.Generator().manual_seed(3)
    table = torch.tensor([[0.7, -1.3], [-0.4, 0.9], [1.6, -0.2], [-1.1, 0.5]])
    with torch.no_grad():
        model.leg_coupling.coupling.copy_(table.cuda())
    assert model.leg_coupling.coupling.requires_grad
    assert any(v is model.leg_coupling.coupling for v in model.trainable_variables())
    B = 777                                     # ragged against the 1024-thread reduction
    idx = torch.randint(0, 4, (B, 1), generator=g)
    x = torch.randn(B, 19, generator=g)
    w_out = torch.randn(B, 4, generator=g)
    w_val = torch.randn(B, generator=g)
    out, _ = model({"obs": (idx.cuda(), x.cuda())}, [], None)
    loss = (out * w_out.cuda()).sum() + (model.value_function() * w_val.cuda()).sum()
    model.theta.grad = None
    model.leg_coupling.coupling.grad = None
    loss.backward()
    th = model.theta.detach().cpu().double().requires_grad_(True)
    tb = table.double().requires_grad_(True)
    lg, vr = O.fcnet_forward(th, x.double(), 4)
    ref = O.leg_coupling(lg, idx, tb)
    ((ref * w_out.double()).sum() + (vr * w_val.double()).sum()).backward()
    assert scaled_err(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    assert scaled_err(model.leg_coupling.coupling.grad.cpu().numpy(), tb.grad.numpy()) < TOL
    assert scaled_err(model.theta.grad.cpu().numpy(), th.grad.numpy()) < 2e-5
    # one plain SGD step moves the table (it was frozen at its initial value before round 2)
    with torch.no_grad():
        model.leg_coupling.coupling -= 1e-3 * model.leg_coupling.coupling.grad
    assert not torch.equal(model.leg_coupling.coupling.detach().cpu(), table)


def test_unsupported_model_configs_fail_loudly():
    from ddrl_b200 import spaces
    from ddrl_b200._lib import DDRLError
    from ddrl_b200.modelv2 import FullyConnectedNetwork_GlorotUniformInitializer as M
    obs, act = spaces.Box(-np.inf, np.inf, (19,), np.float64), spaces.Box(-1.0, 1.0, (2,))
    for bad in ({"fcnet_hiddens": [128, 128]}, {"fcnet_activation": "relu"}, {"no_final_linear": True}):
        with pytest.raises(DDRLError, match="unsupported"):
            M(obs, act, 4, dict(MODEL_CONFIG, **bad), "p")
