"""CPU: host-side mirror of the reference interface — registry names, multi-agent policy mapping, index tables,
config keys, learner-stat finalisation."""
import numpy as np
import pytest

import oracle.ddrl_oracle as O
from ddrl_b200 import policies as Pz
from ddrl_b200.config import PPOConfig, RLLIB_DEFAULTS


def test_registry_names_match_reference_models_init():
    import ddrl_b200.modelv2 as M
    from ddrl_b200.catalog import ModelCatalog
    assert ModelCatalog.get_custom_model("ffn") is M.FullyConnectedNetwork_GlorotUniformInitializer
    assert ModelCatalog.get_custom_model("fc_glorot_uniform_init") is M.FullyConnectedNetwork_GlorotUniformInitializer
    assert ModelCatalog.get_custom_model("gnn") is M.FullyConnectedNetwork_GNN_GlorotUniformInitializer
    assert ModelCatalog.get_custom_model("cup") is M.FullyConnectedNetwork_Coupling_GlorotUniformInitializer
    with pytest.raises(KeyError):
        ModelCatalog.get_custom_model("nope")


def test_variable_layouts_match_oracle_and_checkpoint_order():
    import ddrl_b200.modelv2 as M
    assert M.fcnet_shapes(19, 4) == [(n, tuple(s)) for n, s in O.fcnet_shapes(19, 4)]
    got = [(n.split("/", 1)[1], s) for n, s in M.graphnet_shapes(4)]
    assert got == O.graphnet_shapes(4) + O.graphnet_shapes(1)


ARCH_TABLE = {  # SURVEY.md §8 architecture table
    "QuantrupedMultiEnv_Centralized": (1, 1, 43, 8),
    "QuantrupedMultiEnv_FullyDecentral": (4, 4, 19, 2),
    "QuantrupedMultiEnv_Local": (4, 4, 35, 2),
    "QuantrupedMultiEnv_SingleNeighbor": (4, 4, 27, 2),
    "QuantrupedMultiEnv_SingleDiagonal": (4, 4, 27, 2),
    "QuantrupedMultiEnv_SingleToFront": (4, 4, 27, 2),
    "QuantrupedMultiEnv_TwoSides": (2, 2, 27, 4),
    "QuantrupedMultiEnv_TwoDiags": (2, 2, 27, 4),
    "QuantrupedMultiEnv_SharedDecentral": (1, 4, 19, 2),
}


@pytest.mark.parametrize("scope", list(ARCH_TABLE))
def test_architecture_table(scope):
    P, Ag, D, A = ARCH_TABLE[scope]
    env = Pz.ARCHITECTURES[scope]
    assert len(env.policy_names) == P and len(env.agent_names) == Ag
    for tv in (False, True):
        pol = env.return_policies(use_target_velocity=tv)
        assert list(pol) == env.policy_names
        for _, obs_space, act_space, cfg in pol.values():
            assert obs_space.shape == (D + int(tv),) and act_space.shape == (A,) and cfg == {}
        if scope != "QuantrupedMultiEnv_Centralized":
            tab = env.gather_table(tv)
            assert tab.shape == (Ag, D + int(tv)) and tab.dtype == np.int32 and tab.max() <= 42 + int(tv)
    mc = Pz.multiagent_config(scope)
    assert set(mc) == {"policies", "policy_mapping_fn", "policies_to_train"} and mc["policies_to_train"] == env.policy_names
    for a in env.agent_names:
        assert mc["policy_mapping_fn"](a) in env.policy_names


def test_policy_mapping_and_indices_match_reference():
    FD = Pz.QuantrupedFullyDecentralizedEnv
    assert [FD.policy_mapping_fn(a) for a in ("agent_FL", "agent_HL", "agent_HR", "agent_FR", "agent_FL_7")] == \
        ["policy_FL", "policy_HL", "policy_HR", "policy_FR", "policy_FL"]
    assert FD.policy_mapping_fn("something_else") == "policy_FR"          # the reference's final else branch
    assert FD.obs_indices()["agent_FL"] == O.get_obs_indices(["body", "fl"])
    assert FD.action_indices() == {"agent_FL": [2, 3], "agent_HL": [4, 5], "agent_HR": [6, 7], "agent_FR": [0, 1]}
    L = Pz.Quantruped_Local_Env.obs_indices()
    assert L["agent_FL"] == O.get_obs_indices(["body", "fl", "hl", "fr"]) and len(L["agent_HR"]) == 35
    SD = Pz.Quantruped_LocalSingleDiagonalLeg_Env.obs_indices()
    assert SD["agent_HR"] == SD["agent_FL"] == O.get_obs_indices(["body", "fl", "hr"])
    assert SD["agent_FR"] == SD["agent_HL"]
    TS = Pz.Quantruped_TwoSideControllers_Env
    assert TS.action_indices() == {"agent_LEFT": [2, 3, 4, 5], "agent_RIGHT": [6, 7, 0, 1]}
    assert TS.policy_mapping_fn("agent_LEFT") == "policy_LEFT" and TS.policy_mapping_fn("agent_RIGHT") == "policy_RIGHT"
    TD = Pz.Quantruped_TwoDiagControllers_Env
    assert TD.action_indices() == {"agent_FLHR": [2, 3, 6, 7], "agent_HLFR": [4, 5, 0, 1]}
    assert Pz.Quantruped_Centralized_Env.policy_mapping_fn("central_agent") == "central_policy"
    tv = Pz.QuantrupedFullyDecentralizedEnv.obs_indices(True)["agent_HL"]
    assert tv[11] == 43 and len(tv) == 20                                   # TVel: index 43 follows the body block


def test_graph_env_tables():
    G = Pz.QuantrupedDecentralizedSharedGraphEnv
    np.testing.assert_array_equal(G.create_adj(), O.ring_adjacency().numpy())
    pol = G.return_policies()
    (_, space, act, _), = pol.values()
    idx_space, obs_space, adj_space = space
    assert obs_space.shape == (4, 23) and adj_space.shape == (4, 4) and idx_space.shape == (1,) and act.shape == (2,)
    assert G.policy_mapping_fn("agent_HR") == "leg_policy"


def test_config_from_rllib_dict_uses_published_values():
    c = PPOConfig.from_rllib({"lambda": 0.9, "sgd_minibatch_size": 4096})
    assert (c.gamma, c.lambda_, c.clip_param, c.vf_clip_param, c.vf_loss_coeff, c.kl_coeff, c.kl_target, c.lr,
            c.grad_clip, c.num_sgd_iter, c.sgd_minibatch_size) == (0.99, 0.9, 0.2, 10.0, 0.5, 0.2, 0.01, 3e-4, 0.5, 10, 4096)
    assert RLLIB_DEFAULTS["observation_filter"] == "MeanStdFilter" and RLLIB_DEFAULTS["train_batch_size"] == 16000


def test_finalize_stats_matches_oracle_definitions():
    from ddrl_b200.learner import finalize_stats
    import torch
    torch.manual_seed(0)
    cfg_o, cfg = O.PPOConfig(entropy_coeff=0.01), PPOConfig(entropy_coeff=0.01)
    n, A = 64, 2
    lg = torch.randn(n, 2 * A, dtype=torch.float64)
    act = O.dg_sample(lg, torch.randn(n, A, dtype=torch.float64))
    old = lg + 0.1 * torch.randn_like(lg)
    v, R, vp = torch.randn(n, dtype=torch.float64) * 5, torch.randn(n, dtype=torch.float64) * 5, torch.randn(n, dtype=torch.float64)
    adv = torch.randn(n, dtype=torch.float64)
    _, st = O.ppo_loss_from_outputs(lg, v, act, old, O.dg_logp(old, act), vp, adv, R, 0.3, cfg_o)
    ratio = torch.exp(O.dg_logp(lg, act) - O.dg_logp(old, act))
    surr = torch.minimum(adv * ratio, adv * ratio.clamp(0.8, 1.2))
    vf = torch.maximum((v - R) ** 2, (vp + (v - vp).clamp(-10, 10) - R) ** 2)
    sums = np.array([[[float((-surr).sum()), float(O.dg_kl(old, lg).sum()), float(vf.sum()), float(O.dg_entropy(lg).sum()),
                       float(R.sum()), float((R * R).sum()), float((R - v).sum()), float(((R - v) ** 2).sum())]]])
    out = finalize_stats(sums, np.array([0.3]), cfg, n)[0]
    for k in O.STAT_KEYS:
        assert abs(out[k] - float(st[k])) < 1e-5 * max(1.0, abs(float(st[k]))), k
    assert out["cur_kl_coeff"] == float(np.float32(0.3)) and out["cur_lr"] == float(np.float32(3e-4))


def test_contact_force_tables_match_the_reference_env_constructors():
    """get_contact_force_indices(['body', leg], [1/4, 1]) etc. — the literal calls of the env constructors
    (quantruped_fourDecentralizedController_environments.py:31-36, …twoDecentralized…:66-69, …centralized…:54-56)."""
    import oracle.ddrl_oracle as O
    from ddrl_b200.policies import ARCHITECTURES
    lit = {
        "QuantrupedMultiEnv_FullyDecentral": {f"agent_{L}": (["body", L.lower()], [1. / 4., 1.]) for L in ("FL", "HL", "HR", "FR")},
        "QuantrupedMultiEnv_Local": {f"agent_{L}": (["body", L.lower()], [1. / 4., 1.]) for L in ("FL", "HL", "HR", "FR")},
        "QuantrupedMultiEnv_TwoSides": {"agent_LEFT": (["body", "fl", "hl"], [1. / 2., 1., 1.]),
                                        "agent_RIGHT": (["body", "hr", "fr"], [1. / 2., 1., 1.])},
        "QuantrupedMultiEnv_TwoDiags": {"agent_FLHR": (["body", "fl", "hr"], [1. / 2., 1., 1.]),
                                        "agent_HLFR": (["body", "hl", "fr"], [1. / 2., 1., 1.])},
        "QuantrupedMultiEnv_Centralized": {"central_agent": (None, None)},
    }
    for scope, agents in lit.items():
        env = ARCHITECTURES[scope]
        mine = env.contact_force_indices()
        table = env.contact_table()
        assert table.shape == (len(env.agent_names), 14)
        for i, a in enumerate(env.agent_names):
            idx, w = O.get_contact_force_indices(*agents[a])
            assert list(mine[a][0]) == list(idx)
            assert np.allclose(np.asarray(mine[a][1], dtype=np.float64), np.asarray(w, dtype=np.float64))
            dense = np.zeros(14)
            for j, wt in zip(idx, w):
                dense[j] += np.asarray(wt).reshape(-1)[0]
            assert np.array_equal(table[i], dense)
    assert ARCHITECTURES["QuantrupedMultiEnv_FullyDecentral"].action_table().tolist() == [[2, 3], [4, 5], [6, 7], [0, 1]]


@pytest.mark.parametrize("D,A", [(19, 2), (43, 8)])
@pytest.mark.parametrize("vf_share,free_std", [(True, False), (False, True), (True, True), (False, False)])
def test_fcnet_optional_layouts_map_onto_the_two_branch_kernel_layout(D, A, vf_share, free_std):
    """`vf_share_layers` / `free_log_std` (models/fcnet_glorot_uniform_init.py:30-36,85-113) are served by the kernels'
    fixed two-branch layout through `fcnet_layout_map`: the oracle's variant forward on the model's own variables equals its
    standard forward on the mapped vector, and autograd through the index-select ties the gradients."""
    import torch
    import oracle.ddrl_oracle as O
    from ddrl_b200.modelv2 import fcnet_layout_map, fcnet_variant_shapes
    shapes = fcnet_variant_shapes(D, 2 * A, vf_share, free_std)
    assert shapes == O.fcnet_shapes(D, 2 * A, (64, 64), vf_share, free_std)          # the reference's variable order
    g = torch.Generator().manual_seed(D + A)
    theta = O.fcnet_init(D, 2 * A, g, vf_share_layers=vf_share, free_log_std=free_std, dtype=torch.float64)
    theta = (theta + 0.05 * torch.randn(theta.shape, generator=g, dtype=torch.float64)).requires_grad_(True)
    m = fcnet_layout_map(D, 2 * A, vf_share, free_std)
    assert m.numel() == O.n_params(O.fcnet_shapes(D, 2 * A)) and int(m.max()) <= theta.numel() and int(m.min()) >= 0
    if not (vf_share or free_std):
        assert torch.equal(m, torch.arange(theta.numel()))
    full = torch.cat([theta, theta.new_zeros(1)])[m]
    x = torch.randn(50, D, generator=g, dtype=torch.float64)
    lg_v, v_v = O.fcnet_forward(theta, x, 2 * A, vf_share_layers=vf_share, free_log_std=free_std)
    lg_f, v_f = O.fcnet_forward(full, x, 2 * A)
    assert torch.equal(lg_v, lg_f) and torch.equal(v_v, v_f)
    wl, wv = torch.randn(50, 2 * A, generator=g, dtype=torch.float64), torch.randn(50, generator=g, dtype=torch.float64)
    (g_v,) = torch.autograd.grad((lg_v * wl).sum() + (v_v * wv).sum(), theta, retain_graph=True)
    (g_f,) = torch.autograd.grad((lg_f * wl).sum() + (v_f * wv).sum(), theta)
    assert float((g_v - g_f).abs().max()) <= 1e-12 * float(g_v.abs().max())


def test_config_defaults_equal_every_published_run():
    """`RLLIB_DEFAULTS`, `PPOConfig.from_rllib` and `DEFAULT_MODEL_CONFIG` against the 120 published params.json files
    (tests/golden/make_params_golden.py): every learner-path hyper-parameter has ONE value across all runs."""
    import json
    import os
    from ddrl_b200.config import DEFAULT_MODEL_CONFIG, RLLIB_DEFAULTS, PPOConfig
    from tests.util import GOLDEN
    g = json.load(open(os.path.join(GOLDEN, "published_params.json")))
    assert g["n_runs"] == 120
    for k, v in RLLIB_DEFAULTS.items():
        assert g["top"][k] == [v], k
    c = PPOConfig.from_rllib({})
    for field_, key in (("gamma", "gamma"), ("lambda_", "lambda"), ("clip_param", "clip_param"), ("vf_clip_param", "vf_clip_param"),
                        ("vf_loss_coeff", "vf_loss_coeff"), ("entropy_coeff", "entropy_coeff"), ("kl_coeff", "kl_coeff"),
                        ("kl_target", "kl_target"), ("lr", "lr"), ("grad_clip", "grad_clip"), ("num_sgd_iter", "num_sgd_iter"),
                        ("sgd_minibatch_size", "sgd_minibatch_size")):
        assert [getattr(c, field_)] == g["top"][key], field_
    for k in ("custom_model", "fcnet_hiddens", "fcnet_activation", "free_log_std", "no_final_linear"):
        assert g["model"][k] == [DEFAULT_MODEL_CONFIG[k]], k
    # params.json is dumped before PPO's setup_config copies the TOP-LEVEL vf_share_layers (False in all runs) into the model
    # config; the checkpoints' fc_value_* variables confirm the separate value network
    assert g["model"]["vf_share_layers"] == [True] and g["top"]["vf_share_layers"] == [False]
    assert DEFAULT_MODEL_CONFIG["vf_share_layers"] is False
    assert g["top"]["lr_schedule"] == [None] and g["top"]["entropy_coeff_schedule"] == [None]       # constant lr / entropy coeff
    assert g["env_config"]["ctrl_cost_weight"] == [0.25, 0.5] and g["env_config"]["contact_cost_weight"] == [0.025, 0.05]


def test_modelv2_fcnet_layouts_orchestration_with_oracle_mocked_kernels():
    """tests/host_dryrun_modelv2.py: the ModelV2 GPU tests (default, vf_share_layers, free_log_std) executed on CPU with the two
    FCNet kernels replaced by the oracle — host logic only (own process: it monkeypatches torch)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "host_dryrun_modelv2.py")], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 5


def test_graphnet_learner_orchestration_with_oracle_mocked_kernels():
    """tests/host_dryrun_graphnet.py: one GraphNetLearner iteration against the oracle and learn_on_batch == the SGD phase of
    learn_on_rollout, executed on CPU with every kernel replaced by the oracle — host logic only (own process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "host_dryrun_graphnet.py")], capture_output=True, text=True,
                         timeout=900, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert out.stdout.split() == ["iteration", "ok", "learn_on_batch", "ok"]


def test_finalize_stats_vectorised_equals_the_per_policy_loop():
    """`learner.finalize_stats` works on all policies at once (the GPU idles while the host computes it between two
    iterations); it must return exactly what the per-policy loop it replaced returned — float32 means like RLlib's
    `_averaged`, same summation order."""
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import finalize_stats

    def loop(sums, kl_coeff, cfg, rows):
        n, out = float(rows), []
        for p in range(sums.shape[1]):
            s = sums[:, p, :]
            pol, kl, vf, ent = s[:, 0] / n, s[:, 1] / n, s[:, 2] / n, s[:, 3] / n
            total = pol + kl_coeff[p] * kl + cfg.vf_loss_coeff * vf - cfg.entropy_coeff * ent
            yvar = s[:, 5] / n - (s[:, 4] / n) ** 2
            dvar = s[:, 7] / n - (s[:, 6] / n) ** 2
            with np.errstate(divide="ignore", invalid="ignore"):
                ev = np.maximum(-1.0, 1.0 - dvar / yvar)
            m = lambda a: float(np.mean(a.astype(np.float32)))      # noqa: E731
            out.append({"total_loss": m(total), "policy_loss": m(pol), "vf_loss": m(vf), "kl": m(kl), "entropy": m(ent),
                        "vf_explained_var": m(ev), "cur_kl_coeff": float(np.float32(kl_coeff[p])),
                        "cur_lr": float(np.float32(cfg.lr)), "entropy_coeff": float(cfg.entropy_coeff)})
        return out

    rng = np.random.default_rng(5)
    cfg = PPOConfig(entropy_coeff=0.01)
    for _ in range(40):
        nb, P = int(rng.integers(1, 130)), int(rng.integers(1, 7))
        sums = rng.standard_normal((nb, P, 8)) * 100.0 + 500.0
        sums[..., 5] = np.abs(sums[..., 5]) * 50.0
        kl = rng.random(P) + 0.1
        assert finalize_stats(sums, kl, cfg, 4096.0) == loop(sums, kl, cfg, 4096.0)
