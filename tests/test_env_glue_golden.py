"""CPU: the oracle's env-glue restatements against vectors produced by the REFERENCE'S OWN numpy code
(methods lifted from the reference files and executed unmodified: tests/golden/make_env_glue_golden.py,
tests/golden/make_graph_obs_golden.py).  Bit-exact: float64 numpy on both sides."""
import os

import numpy as np
import pytest

import oracle.ddrl_oracle as O
from ddrl_b200 import policies as P
from tests.util import GOLDEN

GLUE = np.load(os.path.join(GOLDEN, "env_glue.npz"))
GRAPH = np.load(os.path.join(GOLDEN, "graph_obs.npz"))
TAGS = {"four": "QuantrupedMultiEnv_FullyDecentral", "two": "QuantrupedMultiEnv_TwoSides", "one": "QuantrupedMultiEnv_Centralized"}


@pytest.mark.parametrize("tag", list(TAGS))
def test_index_and_weight_tables_equal_the_reference_constructors(tag):
    env = P.ARCHITECTURES[TAGS[tag]]
    assert np.array_equal(env.action_table(), GLUE[f"{tag}/action_idx"])
    assert np.array_equal(env.contact_table(), GLUE[f"{tag}/contact_w"])


@pytest.mark.parametrize("tag", list(TAGS))
@pytest.mark.parametrize("mode", ["per_leg", "per_leg_norm", "global", "global_costs"])
def test_oracle_reward_split_reproduces_the_reference_adaptor(tag, mode):
    env = P.ARCHITECTURES[TAGS[tag]]
    fw, act, cfrc, want = GLUE[f"{tag}/fw"], GLUE[f"{tag}/act"], GLUE[f"{tag}/cfrc"], GLUE[f"{tag}/{mode}"]
    cfi = {a: O.get_contact_force_indices(*(([["body", *env._act_prefixes[a]], [len(env._act_prefixes[a]) / 4.0] + [1.0] * len(env._act_prefixes[a])])
                                             if env._act_prefixes[a] is not None else [])) for a in env.agent_names}
    for s in range(len(fw)):
        r = O.distribute_rewards(fw[s], {a: act[s, i] for i, a in enumerate(env.agent_names)}, cfrc[s], cfi, env.agent_names,
                                 0.25, 0.025, mode)
        assert np.array_equal(np.asarray([r[a] for a in env.agent_names]), want[s]), (s, mode)


@pytest.mark.parametrize("tag", list(TAGS))
def test_oracle_concatenate_actions_reproduces_the_reference(tag):
    env = P.ARCHITECTURES[TAGS[tag]]
    act, want = GLUE[f"{tag}/act"], GLUE[f"{tag}/actions"]
    ai = {a: O.get_action_indices(env._act_prefixes[a]) if env._act_prefixes[a] is not None else list(range(8)) for a in env.agent_names}
    for s in range(len(act)):
        assert np.array_equal(O.concatenate_actions({a: act[s, i] for i, a in enumerate(env.agent_names)}, ai), want[s])


def test_graph_obs_table_is_the_reference_prefix_major_order():
    env = P.ARCHITECTURES["QuantrupedMultiEnv_DecentralShared_Graph"]
    assert np.array_equal(env.gather_table(), GRAPH["table"])
    assert env.leg_angles == O.LEG_ANGLES


def test_oracle_graph_node_features_reproduce_the_reference_env_sequentially():
    """The env filter updates on every call (quantruped_adaptor_multi_environment.py:83-85): replay the 64 steps."""
    obs, want = GRAPH["obs_full"], GRAPH["seq"]
    idx = {a: O.get_obs_indices(["body", a[-2:].lower()]) for a in ["agent_FL", "agent_HL", "agent_HR", "agent_FR"]}
    filt = O.MeanStdFilter((43,))
    for t in range(len(obs)):
        assert np.array_equal(O.graph_distribute_observations(obs[t], filt, idx), want[t]), t
    assert np.array_equal(filt.rs.mean, GRAPH["mean"]) and np.array_equal(filt.rs.std, GRAPH["std"])


def test_oracle_graph_node_features_with_frozen_filter():
    obs, want = GRAPH["obs_full"], GRAPH["frozen"]
    idx = {a: O.get_obs_indices(["body", a[-2:].lower()]) for a in ["agent_FL", "agent_HL", "agent_HR", "agent_FR"]}
    mean, std = GRAPH["mean"], GRAPH["std"]
    norm = lambda x: np.clip((x - mean) / (std + 1e-8), -10.0, 10.0)
    for t in range(len(obs)):
        assert np.array_equal(O.graph_distribute_observations(obs[t], norm, idx), want[t]), t
    q = want[:, :, 19:]                                   # unit body quaternion times a unit yaw quaternion stays unit
    assert np.allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-12)


# ---- a13: the architecture table against the reference's own static env interface ------------------------------------------
import json  # noqa: E402

ARCH = json.load(open(os.path.join(GOLDEN, "architectures.json")))


def _describe(sp):
    from ddrl_b200 import spaces
    if isinstance(sp, spaces.Tuple):
        return ["Tuple", [_describe(s) for s in sp]]
    if isinstance(sp, spaces.MultiDiscrete):
        return ["MultiDiscrete", np.asarray(sp.nvec).astype(int).tolist()]
    return ["Box", list(sp.shape), str(np.dtype(sp.dtype))]


def test_every_policy_scope_of_the_training_script_is_mirrored():
    assert set(ARCH) == set(P.ARCHITECTURES)


@pytest.mark.parametrize("scope", sorted(ARCH))
def test_architecture_table_equals_the_reference_classes(scope):
    """policy / agent names, `policy_mapping_fn` on known and unknown agent ids, and the spaces of `return_policies`
    (flat and target-velocity variants) — from tests/golden/make_arch_golden.py, which executes the reference's static
    class members.  One documented deviation: the fork keys the centralized policy "centr_A_policy" in return_policies
    while its policy_mapping_fn and every published checkpoint say "central_policy"; the published name is kept."""
    ref, env = ARCH[scope], P.ARCHITECTURES[scope]
    assert env.__name__ == ref["class"]
    assert list(env.policy_names) == ref["policy_names"] and list(env.agent_names) == ref["agent_names"]
    for agent_id, pid in ref["mapping"].items():
        assert env.policy_mapping_fn(agent_id) == pid, agent_id
    for tag, tv in (("flat", False), ("tvel", True)):
        got = {pid: {"obs": _describe(spec[1]), "act": _describe(spec[2])}
               for pid, spec in env.return_policies(use_target_velocity=tv).items()}
        want = ref["policies"][tag]
        if scope == "QuantrupedMultiEnv_Centralized":
            want = {"central_policy": want["centr_A_policy"]}
        assert got == want, tag
    assert P.multiagent_config(scope)["policies_to_train"] == ref["policy_names"]


def test_leg_transform_action_scale():
    s = P.ARCHITECTURES["QuantrupedMultiEnv_SharedDecentralLegTransforms"].action_scale()
    assert s.tolist() == [1.0, -1.0, 1.0, 1.0, 1.0, 1.0, 1.0, -1.0]       # fr_knee, hr_knee (ACTION_FIELDS order)


@pytest.mark.parametrize("scope", sorted(ARCH))
def test_index_tables_equal_what_the_reference_constructors_build(scope):
    """obs / action / contact-force index tables of every agent of every scope, from the reference CONSTRUCTORS executed on a
    simulation-free root class (tests/golden/make_arch_golden.py) — includes the SingleDiagonal quirk (HR and FR reuse the
    FL / HL lists) and the prefix-major ordering."""
    ref, env = ARCH[scope]["tables"], P.ARCHITECTURES[scope]
    assert {a: list(map(int, v)) for a, v in env.obs_indices().items()} == ref["obs"]
    assert {a: list(map(int, v)) for a, v in env.action_indices().items()} == ref["act"]
    got = env.contact_force_indices()
    assert set(got) == set(ref["contact"])
    for a, (idx, w) in got.items():
        assert [int(i) for i in idx] == ref["contact"][a][0], a
        assert [float(np.asarray(x).reshape(-1)[0]) for x in w] == ref["contact"][a][1], a
    if ref["obs"] and scope not in ("QuantrupedMultiEnv_Centralized",):
        assert env.gather_table().tolist() == [ref["obs"][a] for a in env.agent_names]
    # target-velocity variant: the 44-field list QuAntrupedTVelEnv declares (index 43 joins the body block)
    tv = ARCH[scope]["tables_tvel"]["obs"]
    assert {a: list(map(int, v)) for a, v in env.obs_indices(use_target_velocity=True).items()} == tv
    assert all(len(v) == len(ref["obs"][a]) + 1 and 43 in v for a, v in tv.items())


# ---- one whole env step as the reference wires it (constructors + step() executed, simulation stubbed) ----------------------
STEP = np.load(os.path.join(GOLDEN, "env_step.npz"))
MODES = {("distribute_per_leg_reward", False): "per_leg", ("distribute_per_leg_reward", True): "per_leg_norm",
         ("distribute_global_reward", False): "global", ("distribute_global_reward", True): "global"}
CFGS = {"per_leg": {}, "norm": {"norm_reward": True}, "global": {"global_reward": True}}


@pytest.mark.parametrize("scope", sorted(ARCH))
@pytest.mark.parametrize("cfg", sorted(CFGS))
def test_env_step_rewards_actions_and_observations(scope, cfg):
    env = P.ARCHITECTURES[scope]
    agents = list(env.agent_names)
    wired = MODES[(str(STEP[f"{scope}/{cfg}/reward_fn"]), bool(STEP[f"{scope}/{cfg}/normalize_rewards"]))]
    assert P.reward_mode(CFGS[cfg]) == wired          # also for the GlobalCost scope: its override is shadowed as shipped
    acts, fw, cfrc, obs_full = STEP[f"actions/{scope}"], STEP["fw"], STEP["cfrc"], STEP["obs_full"]
    cfi, ai = env.contact_force_indices(), env.action_indices()
    filt = O.MeanStdFilter((43,))                      # the env-side singleton: one update per step, before the gather
    node_filt = O.MeanStdFilter((4, 19))               # Decentral_Graph normalises the gathered node matrix instead
    for t in range(len(fw)):
        ad = {a: acts[t, i] for i, a in enumerate(agents)}
        want_act = O.concatenate_actions(ad, {a: (ai[a] if env._act_prefixes[a] is not None else list(range(8))) for a in agents})
        if scope.endswith("LegTransforms"):
            want_act = want_act * env.action_scale()
        assert np.array_equal(STEP[f"{scope}/{cfg}/sim_action"][t], want_act), t
        r = O.distribute_rewards(fw[t], ad, cfrc[t], cfi, agents, 0.25, 0.025, wired)
        assert np.array_equal(STEP[f"{scope}/{cfg}/rew"][t], np.asarray([r[a] for a in agents])), t
        got = STEP[f"{scope}/{cfg}/obs"][t]
        oi = env.obs_indices()
        if scope == "QuantrupedMultiEnv_DecentralShared_Graph":
            want = O.graph_distribute_observations(obs_full[t], filt, oi)
            assert all(np.array_equal(got[i], want) for i in range(len(agents))), t
        elif scope == "QuantrupedMultiEnv_Decentral_Graph":
            want = node_filt(np.stack([obs_full[t][oi[a]] for a in agents]))
            assert all(np.array_equal(got[i], want) for i in range(len(agents))), t
        else:
            normed = filt(obs_full[t])
            for i, a in enumerate(agents):
                idx = oi[a] if env._obs_prefixes[a] is not None else list(range(43))
                assert np.array_equal(got[i], normed[idx]), (t, a)
