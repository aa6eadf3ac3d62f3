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
