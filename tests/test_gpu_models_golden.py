"""GPU: the CUDA model kernels directly against vectors produced by the reference's own model code
(tests/golden/make_models_golden.py: `/root/reference/models/*.py` executed on a numpy stand-in for TensorFlow) — the same
kernels the other GPU tests compare with the oracle, here on the golden inputs and weights (rounded to float32)."""
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
G = np.load(os.path.join(GOLDEN, "models.npz"))


def dev(k, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(G[k].astype(dtype))).cuda()


def test_graphnet_wrapper_kernel_equals_reference_model():
    from ddrl_b200 import kernels as K
    lg, v = K.graphnet_forward(dev("wrap/theta"), dev("gn/idx", np.int32).reshape(-1), dev("gn/state"), dev("gn/adj"), 2)
    assert scaled_err(lg.cpu().numpy(), G["wrap/logits"]) < TOL
    assert scaled_err(v.cpu().numpy(), G["wrap/value"]) < TOL


def test_fcnet_kernels_equal_reference_model():
    from ddrl_b200 import kernels as K
    theta, x = dev("fc/default/theta").reshape(1, -1), dev("fc/x").reshape(1, 13, 19)
    out = K.fcnet_forward(theta, x, 2)
    assert scaled_err(out["logits"][0].cpu().numpy(), G["fc/default/logits"]) < TOL
    assert scaled_err(out["value"][0].cpu().numpy(), G["fc/default/value"]) < TOL
    tc = K.fcnet_forward_tc(K.fcnet_tc_pack(theta, 19, 2), x, 2)          # tensor-core inference forward
    assert int(tc["status"].item()) == 0
    assert scaled_err(tc["logits"][0].cpu().numpy(), G["fc/default/logits"]) < TOL
    assert scaled_err(tc["value"][0].cpu().numpy(), G["fc/default/value"]) < TOL


@pytest.mark.parametrize("adj", ["ring", "rand"])
def test_graph_layer_kernels_equal_reference_layers(adj):
    from ddrl_b200 import kernels as K
    x, a = dev("layers/x"), dev(f"layers/adj_{adj}")
    y = K.mpnn2_forward(x, a, dev("mpnn2/W_msg"), dev("mpnn2/W_upd"), dev("mpnn2/b"), "tanh")
    assert scaled_err(y.cpu().numpy(), G[f"mpnn2/y_{adj}"]) < TOL
    y = K.gat1_forward(x, a, dev("gat1/W_pre"), dev("gat1/w_att"), dev("gat1/b"), "tanh")
    assert scaled_err(y.cpu().numpy(), G[f"gat1/y_{adj}"]) < TOL


def test_gcn_norms_and_coupling_kernels_equal_reference():
    from ddrl_b200 import kernels as K
    y = K.gcn_forward(dev("layers/x"), dev("gcn/adj"), dev("gcn/W"), dev("gcn/b"), "tanh")
    assert scaled_err(y.cpu().numpy(), G["gcn/y"]) < TOL
    assert scaled_err(K.symm_norm(dev("ops/adj")).cpu().numpy(), G["ops/symm_norm"]) < TOL
    for c in range(3):       # one column per call (the softmax is per column; the GPU suite exercises C = 1)
        sm = K.segment_softmax(dev("ops/seg_data")[:, c:c + 1].contiguous(), dev("ops/seg_ids", np.int32), 5)
        assert scaled_err(sm.cpu().numpy(), G["ops/segment_softmax"][:, c:c + 1]) < TOL
    out = K.leg_coupling_(dev("cup/logits_in"), dev("cup/node_id", np.int32).reshape(-1), dev("cup/coupling"))
    assert np.array_equal(out.cpu().numpy(), G["cup/layer_out"].astype(np.float32))


S = np.load(os.path.join(GOLDEN, "env_step.npz"))
WIRED = {("distribute_per_leg_reward", False): "per_leg", ("distribute_per_leg_reward", True): "per_leg_norm",
         ("distribute_global_reward", False): "global", ("distribute_global_reward", True): "global"}


@pytest.mark.parametrize("scope", ["QuantrupedMultiEnv_FullyDecentral", "QuantrupedMultiEnv_TwoDiags", "QuantrupedMultiEnv_Centralized",
                                   "QuantrupedMultiEnv_Local", "QuantrupedMultiEnv_SharedDecentralLegTransforms",
                                   "QuantrupedMultiEnv_FullyDecentralGlobalCost"])
def test_env_step_kernels_equal_the_reference_step(scope):
    """ddrl_reward_split / ddrl_concat_actions on the inputs of the reference's own env step() (run with a stubbed simulator,
    tests/golden/make_env_step_golden.py), for the three env_configs the training script can select."""
    from ddrl_b200 import kernels as K
    from ddrl_b200 import policies as P
    env = P.ARCHITECTURES[scope]
    to = lambda a, dt=np.float32: torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(dt))).cuda()      # noqa: E731
    act = S[f"actions/{scope}"]
    for cfg, env_config in (("per_leg", {}), ("norm", {"norm_reward": True}), ("global", {"global_reward": True})):
        mode = WIRED[(str(S[f"{scope}/{cfg}/reward_fn"]), bool(S[f"{scope}/{cfg}/normalize_rewards"]))]
        assert P.reward_mode(env_config) == mode
        rew = K.reward_split(to(S["fw"]), to(act), to(S["cfrc"], np.float64), to(env.contact_table(), np.float64), 0.25, 0.025,
                             mode).cpu().numpy()
        want = S[f"{scope}/{cfg}/rew"]
        assert np.abs(rew - want).max() < 2e-6 * max(1.0, np.abs(want).max()), cfg
    full = K.concat_actions(to(act), to(env.action_table(), np.int32))
    if scope.endswith("LegTransforms"):
        full = full * to(env.action_scale())
    assert np.array_equal(full.cpu().numpy(), S[f"{scope}/per_leg/sim_action"].astype(np.float32))
