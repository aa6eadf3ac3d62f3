"""Multi-agent policy mapping of the reference's control architectures, without the MuJoCo environments.

Each class mirrors the static interface RLlib consumes from the reference's env classes
(``policy_names``, ``agent_names``, ``policy_mapping_fn(agent_id)``, ``return_policies(use_target_velocity)``,
train_experiment_1_architecture_on_flat.py:141-151) plus the index tables the batched obs-gather needs
(``obs_indices`` / ``action_indices``, built with the prefix-major rule of
simulation_envs/quantruped_v3.py:282-317).  ``ARCHITECTURES`` is keyed by the ``--policy_scope`` names
(train_experiment_1_architecture_on_flat.py:63-90)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import spaces

# simulation_envs/quantruped_v3.py:68-102
OBS_FIELDS = [
    "body_height", "body_qpos_x", "body_qpos_y", "body_qpos_z", "body_qpos_w",
    "fl_hip", "fl_knee", "hl_hip", "hl_knee", "hr_hip", "hr_knee", "fr_hip", "fr_knee",
    "body_vel_x", "body_vel_y", "body_vel_z", "body_rot_vel_x", "body_rot_vel_y", "body_rot_vel_z",
    "fl_hip_vel", "fl_knee_vel", "hl_hip_vel", "hl_knee_vel", "hr_hip_vel", "hr_knee_vel", "fr_hip_vel", "fr_knee_vel",
    "fl_hip_pforce", "fl_knee_pforce", "hl_hip_pforce", "hl_knee_pforce",
    "hr_hip_pforce", "hr_knee_pforce", "fr_hip_pforce", "fr_knee_pforce",
    "fr_hip_hist_ctrl", "fr_knee_vel_hist_ctrl", "fl_hip_hist_ctrl", "fl_knee_vel_hist_ctrl",
    "hl_hip_hist_ctrl", "hl_knee_vel_hist_ctrl", "hr_hip_hist_ctrl", "hr_knee_vel_hist_ctrl",
]
ACTION_FIELDS = ["fr_hip", "fr_knee", "fl_hip", "fl_knee", "hl_hip", "hl_knee", "hr_hip", "hr_knee"]
# simulation_envs/quantruped_v3.py:105-112 — rows of sim.data.cfrc_ext ([14 bodies][6])
CONTACT_FORCE_FIELDS = ["body_floor", "body", "fl_hip", "fl_leg", "fl_foot", "hl_hip", "hl_leg", "hl_foot",
                        "hr_hip", "hr_leg", "hr_foot", "fr_hip", "fr_leg", "fr_foot"]
TVEL_FIELD = "body_target_x_vel"   # appended as index 43 by QuAntrupedTVelEnv.set_target_velocity (quantruped_v3.py:394-400)


def get_obs_indices(prefixes: Sequence[str] = None, use_target_velocity: bool = False) -> List[int]:
    fields = OBS_FIELDS + ([TVEL_FIELD] if use_target_velocity else [])
    if prefixes is None:
        return list(range(len(fields)))
    out: List[int] = []
    for p in prefixes:
        out.extend(i for i, f in enumerate(fields) if f.startswith(p))
    return out


def get_action_indices(prefixes: Sequence[str] = None) -> List[int]:
    if prefixes is None:
        return list(range(len(ACTION_FIELDS)))
    out: List[int] = []
    for p in prefixes:
        out.extend(i for i, f in enumerate(ACTION_FIELDS) if f.startswith(p))
    return out


def get_contact_force_indices(prefixes: Sequence[str] = None, weights: Sequence[float] = None):
    """quantruped_v3.py:319-341: (indices, weights [k][1]) of the cfrc_ext rows whose field name starts with a prefix."""
    if prefixes is None:
        n = len(CONTACT_FORCE_FIELDS)
        return list(range(n)), [[1.0]] * n
    if weights is None:
        weights = [1.0] * len(prefixes)
    idx: List[int] = []
    w: List[List[float]] = []
    for pfx, wt in zip(prefixes, weights):
        hit = [i for i, f in enumerate(CONTACT_FORCE_FIELDS) if f.startswith(pfx)]
        idx.extend(hit)
        w.extend([[wt]] * len(hit))
    return idx, w


class _Arch:
    """Base: subclasses set policy_names, agent_names, _obs_prefixes, _act_prefixes, _agent_policy."""
    policy_names: List[str] = []
    agent_names: List[str] = []
    _obs_prefixes: Dict[str, Sequence[str]] = {}
    _act_prefixes: Dict[str, Sequence[str]] = {}
    _agent_policy: Dict[str, str] = {}
    model = "fc_glorot_uniform_init"

    @classmethod
    def policy_mapping_fn(cls, agent_id: str) -> str:
        # the reference matches by prefix and falls through to the last policy (e.g. fourDecentralized…:38-48)
        for agent, pol in cls._agent_policy.items():
            if agent_id.startswith(agent):
                return pol
        return cls.policy_names[-1]

    @classmethod
    def obs_indices(cls, use_target_velocity: bool = False) -> Dict[str, List[int]]:
        return {a: get_obs_indices(cls._obs_prefixes[a], use_target_velocity) for a in cls.agent_names}

    @classmethod
    def action_indices(cls) -> Dict[str, List[int]]:
        return {a: get_action_indices(cls._act_prefixes[a]) for a in cls.agent_names}

    @classmethod
    def obs_dim(cls, use_target_velocity: bool = False) -> int:
        return len(cls.obs_indices(use_target_velocity)[cls.agent_names[0]])

    @classmethod
    def act_dim(cls) -> int:
        return len(cls.action_indices()[cls.agent_names[0]])

    @classmethod
    def return_policies(cls, use_target_velocity: bool = False):
        obs_space = spaces.Box(-np.inf, np.inf, (cls.obs_dim(use_target_velocity),), np.float64)
        act_space = spaces.Box(-1.0, 1.0, (cls.act_dim(),))
        return {p: (None, obs_space, act_space, {}) for p in cls.policy_names}

    @classmethod
    def agents_per_policy(cls) -> int:
        return len(cls.agent_names) // len(cls.policy_names)

    @classmethod
    def contact_force_indices(cls):
        """Per agent (indices, weights) like the env constructors build them: the torso rows ('body' matches body_floor
        and body) weighted by 1 / #agents-sharing-them, the agent's own legs by 1
        (e.g. quantruped_fourDecentralizedController_environments.py:31-36, …twoDecentralized…:66-69); the centralized
        controller takes every row with weight 1 (quantruped_centralizedController_environment.py:54-56)."""
        out = {}
        for a in cls.agent_names:
            legs = cls._act_prefixes[a]
            if legs is None:
                out[a] = get_contact_force_indices()
            else:
                out[a] = get_contact_force_indices(["body", *legs], [len(legs) / 4.0] + [1.0] * len(legs))
        return out

    @classmethod
    def contact_table(cls) -> np.ndarray:
        """float64 [n_agents, 14] dense weights of the cfrc_ext rows (0 = row not used by the agent)."""
        t = np.zeros((len(cls.agent_names), len(CONTACT_FORCE_FIELDS)), dtype=np.float64)
        for i, a in enumerate(cls.agent_names):
            idx, w = cls.contact_force_indices()[a]
            for j, wt in zip(idx, w):
                t[i, j] += wt[0]
        return t

    @classmethod
    def action_table(cls) -> np.ndarray:
        """int32 [n_agents, A]: position of each agent action in the 8-dim env action (concatenate_actions)."""
        ai = cls.action_indices()
        return np.asarray([ai[a] for a in cls.agent_names], dtype=np.int32)

    @classmethod
    def gather_table(cls, use_target_velocity: bool = False) -> np.ndarray:
        """int32 [n_agents, D] index table for the device obs gather, agents in ``agent_names`` order."""
        oi = cls.obs_indices(use_target_velocity)
        return np.asarray([oi[a] for a in cls.agent_names], dtype=np.int32)


_LEGS = ("FL", "HL", "HR", "FR")
_leg = {"FL": "fl", "HL": "hl", "HR": "hr", "FR": "fr"}


def _four(name, extra: Dict[str, Sequence[str]]):
    agents = [f"agent_{l}" for l in _LEGS]
    return type(name, (_Arch,), dict(
        policy_names=[f"policy_{l}" for l in _LEGS], agent_names=agents,
        _obs_prefixes={f"agent_{l}": ["body", _leg[l], *extra[l]] for l in _LEGS},
        _act_prefixes={f"agent_{l}": [_leg[l]] for l in _LEGS},
        _agent_policy={f"agent_{l}": f"policy_{l}" for l in _LEGS}))


class Quantruped_Centralized_Env(_Arch):
    """quantruped_centralizedController_environment.py:55-74.  Published checkpoints use ``central_policy``
    (the fork's return_policies keys it ``centr_A_policy``, SURVEY.md §2.3 — the published name is kept)."""
    policy_names = ["central_policy"]
    agent_names = ["central_agent"]
    _obs_prefixes = {"central_agent": None}
    _act_prefixes = {"central_agent": None}
    _agent_policy = {"central_agent": "central_policy"}

    @classmethod
    def return_policies(cls, use_target_velocity: bool = False):
        obs_space = spaces.Box(-np.inf, np.inf, (43 + int(use_target_velocity),), np.float64)
        return {"central_policy": (None, obs_space, spaces.Box(-1.0, 1.0, (8,)), {})}


# quantruped_fourDecentralizedController_environments.py:168-488
QuantrupedFullyDecentralizedEnv = _four("QuantrupedFullyDecentralizedEnv", {l: [] for l in _LEGS})
Quantruped_LocalSingleNeighboringLeg_Env = _four(
    "Quantruped_LocalSingleNeighboringLeg_Env", {"FL": ["hl"], "HL": ["hr"], "HR": ["fr"], "FR": ["fl"]})
Quantruped_LocalSingleDiagonalLeg_Env = _four(
    "Quantruped_LocalSingleDiagonalLeg_Env", {"FL": ["hr"], "HL": ["fr"], "HR": ["hr"], "FR": ["fr"]})
Quantruped_LocalSingleToFront_Env = _four(
    "Quantruped_LocalSingleToFront_Env", {"FL": ["hl"], "HL": ["hr"], "HR": ["hl"], "FR": ["hr"]})
Quantruped_Local_Env = _four(
    "Quantruped_Local_Env", {"FL": ["hl", "fr"], "HL": ["hr", "fl"], "HR": ["fr", "hl"], "FR": ["fl", "hr"]})
# SingleDiagonal: HR and FR reuse the FL / HL index lists verbatim (…fourDecentralized…:337-339)
Quantruped_LocalSingleDiagonalLeg_Env._obs_prefixes["agent_HR"] = ["body", "fl", "hr"]
Quantruped_LocalSingleDiagonalLeg_Env._obs_prefixes["agent_FR"] = ["body", "hl", "fr"]


class Quantruped_TwoSideControllers_Env(_Arch):
    """quantruped_twoDecentralizedController_environments.py:6-90"""
    policy_names = ["policy_LEFT", "policy_RIGHT"]
    agent_names = ["agent_LEFT", "agent_RIGHT"]
    _obs_prefixes = {"agent_LEFT": ["body", "fl", "hl"], "agent_RIGHT": ["body", "hr", "fr"]}
    _act_prefixes = {"agent_LEFT": ["fl", "hl"], "agent_RIGHT": ["hr", "fr"]}
    _agent_policy = {"agent_LEFT": "policy_LEFT", "agent_RIGHT": "policy_RIGHT"}


class Quantruped_TwoDiagControllers_Env(_Arch):
    """quantruped_twoDecentralizedController_environments.py:92-171"""
    policy_names = ["policy_FLHR", "policy_HLFR"]
    agent_names = ["agent_FLHR", "agent_HLFR"]
    _obs_prefixes = {"agent_FLHR": ["body", "fl", "hr"], "agent_HLFR": ["body", "hl", "fr"]}
    _act_prefixes = {"agent_FLHR": ["fl", "hr"], "agent_HLFR": ["hl", "fr"]}
    _agent_policy = {"agent_FLHR": "policy_FLHR", "agent_HLFR": "policy_HLFR"}


class QuantrupedSingleDecentralizedEnv(_Arch):
    """One shared FCNet policy for the four legs (quantruped_singleDecentralizedController_environments.py:21-59)."""
    policy_names = ["policy_legs"]
    agent_names = [f"agent_{l}" for l in _LEGS]
    _obs_prefixes = {f"agent_{l}": ["body", _leg[l]] for l in _LEGS}
    _act_prefixes = {f"agent_{l}": [_leg[l]] for l in _LEGS}
    _agent_policy = {f"agent_{l}": "policy_legs" for l in _LEGS}


class QuantrupedDecentralizedSharedGraphEnv(_Arch):
    """One shared GraphNet policy; obs = Tuple(node_idx[1], obs[4, 19+4], adj[4,4])
    (quantruped_GraphDecentralizedController_environments.py:122-245)."""
    policy_names = ["leg_policy"]
    agent_names = [f"agent_{l}" for l in _LEGS]
    _obs_prefixes = {f"agent_{l}": ["body", _leg[l]] for l in _LEGS}
    _act_prefixes = {f"agent_{l}": [_leg[l]] for l in _LEGS}
    _agent_policy = {f"agent_{l}": "leg_policy" for l in _LEGS}
    model = "gnn"
    leg_angles = {"agent_FL": 45.0, "agent_HL": 135.0, "agent_HR": -135.0, "agent_FR": -45.0}

    @staticmethod
    def create_edge_index():
        n = {a: i for i, a in enumerate(QuantrupedDecentralizedSharedGraphEnv.agent_names)}
        ring = [("agent_FL", "agent_HL"), ("agent_HL", "agent_HR"), ("agent_HR", "agent_FR"), ("agent_FR", "agent_FL")]
        return [[n[s], n[r]] for s, r in ring] + [[n[r], n[s]] for s, r in ring]

    @classmethod
    def leg_quat_table(cls) -> np.ndarray:
        """float64 [4, 2] = {sin, cos} of HALF the leg's mounting angle: the (z, w) components of the yaw quaternion
        `leg_encoding_ego` multiplies the body orientation with (…GraphDecentralized…:145-161)."""
        half = np.deg2rad(np.asarray([cls.leg_angles[a] for a in cls.agent_names]) / 2.0)
        return np.stack((np.sin(half), np.cos(half)), axis=1)

    @classmethod
    def create_adj(cls) -> np.ndarray:
        adj = np.zeros([4, 4], dtype=np.float64)
        adj[(*np.transpose(cls.create_edge_index()),)] = 1.0
        return adj

    @classmethod
    def return_policies(cls, use_target_velocity: bool = False):
        n_dims = 19 + int(use_target_velocity) + 2 + 2
        graph_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (4, n_dims), np.float64),
                                    spaces.MultiDiscrete(np.ones([4, 4]) * 2)])
        return {"leg_policy": (None, graph_space, spaces.Box(-1.0, 1.0, (2,)), {})}


class QuantrupedFullyDecentralizedGlobalCostEnv(QuantrupedFullyDecentralizedEnv):
    """Per-leg controllers meant to share the control-cost term
    (quantruped_fourDecentralizedController_GlobalCosts_environments.py:6-112): same observation / action routing and
    policy mapping as FullyDecentral.  Its `distribute_reward` override (mode "global_costs" of `ddrl_reward_split`) is
    SHADOWED as shipped: the root constructor binds `self.distribute_reward` as an instance attribute
    (quantruped_adaptor_multi_environment.py:52-60), so this scope pays the per-leg reward like every other one
    (`reward_mode`; tests/golden/make_env_step_golden.py runs the reference's step() to show it)."""


class QuantrupedDecentralizedGraphEnv(_Arch):
    """Four per-leg policies over Tuple(node_idx[1], obs[4, 19], adj[4, 4])
    (quantruped_GraphDecentralizedController_environments.py:37-120).  The node width 19 has no leg-encoding columns, so
    it does not fit GraphNet's `state[..., -4:]` split (models/graph_net.py:33) — the table is mirrored, the model is not."""
    policy_names = [f"policy_{l}" for l in _LEGS]
    agent_names = [f"agent_{l}" for l in _LEGS]
    _obs_prefixes = {f"agent_{l}": ["body", _leg[l]] for l in _LEGS}
    _act_prefixes = {f"agent_{l}": [_leg[l]] for l in _LEGS}
    _agent_policy = {f"agent_{l}": f"policy_{l}" for l in _LEGS}
    model = "gnn"
    create_edge_index = staticmethod(QuantrupedDecentralizedSharedGraphEnv.create_edge_index)
    create_adj = classmethod(lambda cls: QuantrupedDecentralizedSharedGraphEnv.create_adj())

    @classmethod
    def return_policies(cls, use_target_velocity: bool = False):
        n_dims = 19 + int(use_target_velocity)
        graph_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (4, n_dims), np.float64),
                                    spaces.MultiDiscrete(np.ones([4, 4]) * 2)])
        return {p: (None, graph_space, spaces.Box(-1.0, 1.0, (2,)), {}) for p in cls.policy_names}


class QuantrupedSingleDecentralizedLegIDEnv(QuantrupedSingleDecentralizedEnv):
    """Shared per-leg policy that also receives the leg index: obs = Tuple(node_idx[1], obs[19]) — the input of the
    LegCoupling model "cup" (quantruped_singleDecentralizedController_environments.py:66-115)."""
    model = "cup"

    @classmethod
    def return_policies(cls, use_target_velocity: bool = False):
        obs_space = spaces.Tuple([spaces.MultiDiscrete([4]),
                                  spaces.Box(-np.inf, np.inf, (19 + int(use_target_velocity),), np.float64)])
        return {"policy_legs": (None, obs_space, spaces.Box(-1.0, 1.0, (2,)), {})}


class QuantrupedSingleDecentralizedLegTransforms(QuantrupedSingleDecentralizedEnv):
    """Shared per-leg policy whose knee actions of the right-hand legs are mirrored before they reach the simulator
    (quantruped_singleDecentralizedController_environments.py:117-148): `concatenate_actions(...) * action_scale`."""

    @classmethod
    def action_scale(cls) -> np.ndarray:
        """float32 [8] in the env's action order (ACTION_FIELDS): -1 for fr_knee and hr_knee, 1 elsewhere."""
        scale = np.ones(len(ACTION_FIELDS), dtype=np.float32)
        for pfx in ("fr_knee", "hr_knee"):
            scale[get_action_indices([pfx])] = -1.0
        return scale


ARCHITECTURES = {
    "QuantrupedMultiEnv_Centralized": Quantruped_Centralized_Env,
    "QuantrupedMultiEnv_FullyDecentral": QuantrupedFullyDecentralizedEnv,
    "QuantrupedMultiEnv_Local": Quantruped_Local_Env,
    "QuantrupedMultiEnv_SingleNeighbor": Quantruped_LocalSingleNeighboringLeg_Env,
    "QuantrupedMultiEnv_SingleDiagonal": Quantruped_LocalSingleDiagonalLeg_Env,
    "QuantrupedMultiEnv_SingleToFront": Quantruped_LocalSingleToFront_Env,
    "QuantrupedMultiEnv_TwoSides": Quantruped_TwoSideControllers_Env,
    "QuantrupedMultiEnv_TwoDiags": Quantruped_TwoDiagControllers_Env,
    "QuantrupedMultiEnv_SharedDecentral": QuantrupedSingleDecentralizedEnv,
    "QuantrupedMultiEnv_DecentralShared_Graph": QuantrupedDecentralizedSharedGraphEnv,
    "QuantrupedMultiEnv_Decentral_Graph": QuantrupedDecentralizedGraphEnv,
    "QuantrupedMultiEnv_FullyDecentralGlobalCost": QuantrupedFullyDecentralizedGlobalCostEnv,
    "QuantrupedMultiEnv_SharedDecentralLegID": QuantrupedSingleDecentralizedLegIDEnv,
    "QuantrupedMultiEnv_SharedDecentralLegTransforms": QuantrupedSingleDecentralizedLegTransforms,
}


def reward_mode(env_config: Optional[dict] = None) -> str:
    """Mode of `ddrl_reward_split` the reference's adaptor is wired to for an env_config — the same for every scope
    (quantruped_adaptor_multi_environment.py:52-62,173-203): 'global' with `global_reward`, else the per-leg split,
    normalised ('per_leg_norm') with `norm_reward`."""
    cfg = env_config or {}
    if cfg.get("global_reward", False):
        return "global"
    return "per_leg_norm" if cfg.get("norm_reward", False) else "per_leg"


def multiagent_config(policy_scope: str, use_target_velocity: bool = False) -> dict:
    """The ``config["multiagent"]`` dict of train_experiment_1_architecture_on_flat.py:145-151."""
    env = ARCHITECTURES[policy_scope]
    return {"policies": env.return_policies(use_target_velocity=use_target_velocity),
            "policy_mapping_fn": env.policy_mapping_fn, "policies_to_train": list(env.policy_names)}
