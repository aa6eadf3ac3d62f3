"""CPU ORACLE — test infrastructure, NOT product code.

A CPU restatement of the DDRL learner hot path (reference: LucaHermes/ddrl + the Ray/RLlib 1.0.1
math its launch scripts select).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file.  Nothing under ``ddrl_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

Parity status
-------------
* The reference ships NO tests.  ``models/*.py`` cannot be imported here
  (tensorflow / ray absent), and the PPO arithmetic lives in the un-vendored dependency
  ``ray[rllib]==1.0.1`` (README.md:47; checkpoints: 1.0.0 / 1.0.1 / 1.0.1.post1) and
  ``tensorflow==2.3.1`` (README.md:51).
* What pins this oracle (tests/test_oracle_golden.py): the 120 shipped checkpoints under
  ``Results/`` — variable names/shapes/order, TF1-Adam slot layout, MeanStdFilter state schema,
  the loss-composition identity on all 360 policies (3e-7 rel), ``cur_kl_coeff`` in 0.2*1.5^k,
  vf_loss >> vf_clip_param (PPO2-style value clipping), parameter-count CSV.
* Pinned by the reference's OWN numpy code executed here (tests/test_env_glue_golden.py; the pure-numpy
  methods are lifted from the reference files' ASTs by tests/golden/make_{env_glue,graph_obs,curriculum}_golden.py
  and run unmodified on stub objects): get_obs_indices / get_action_indices / get_contact_force_indices,
  the reward / cost split in its four variants, concatenate_actions, the shared-graph env's
  distribute_observations / leg_encoding_ego / quaternion_multiply — bit for bit.
* Pinned by the reference's own MODEL code executed on a numpy stand-in for TensorFlow
  (tests/test_models_golden.py; tests/golden/make_models_golden.py + tf_shim.py import /root/reference/models
  unmodified): graph_ops, GCN / MPNN / MPNN2 / GAT1, GraphNet and its actor / critic wrapper, the FCNet in
  all four layouts, LegCoupling — to 1e-12.  (Pins the reference's composition of ops, not TF's kernels.)
* Everything that lives inside RLlib / TF rather than in the reference repo (GAE recursion,
  standardisation eps 1e-4, filter eps 1e-8, Adam eps placement, minibatch slicing) is a restatement with
  no reference-run vector behind it beyond the checkpoint pins:  **parity unpinned** for those functions
  (DESIGN.md says the same).

Every function cites the reference file:line it follows (paths relative to /root/reference) or the
RLlib 1.0.1 module it restates.  dtype is a parameter: float64 is the ground truth the CUDA FP32
path is compared with, float32 is the "torch-CPU twin" that is timed as the CPU baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)

# --------------------------------------------------------------------------------------------
# a2  GlorotUniformScaled                          models/glorot_uniform_scaled_initializer.py:3-22
# --------------------------------------------------------------------------------------------


def glorot_limit(fan_in: int, fan_out: int, scale: float) -> float:
    """VarianceScaling(scale, 'fan_avg', 'uniform'): limit = sqrt(3*scale/((fan_in+fan_out)/2))."""
    return math.sqrt(6.0 * scale / (fan_in + fan_out))


def glorot_uniform_scaled(fan_in: int, fan_out: int, scale: float, gen: torch.Generator,
                          dtype=torch.float32) -> torch.Tensor:
    lim = glorot_limit(fan_in, fan_out, scale)
    return (torch.rand(fan_in, fan_out, generator=gen, dtype=torch.float64) * 2.0 - 1.0).mul_(lim).to(dtype)


# --------------------------------------------------------------------------------------------
# a1  FCNet                                         models/fcnet_glorot_uniform_init.py:17-125
# --------------------------------------------------------------------------------------------

FC_VAR_ORDER = ("fc_1", "fc_value_1", "fc_2", "fc_value_2", "fc_out", "value_out")


def fcnet_shapes(D: int, num_outputs: int, hiddens: Sequence[int] = (64, 64),
                 vf_share_layers: bool = False, free_log_std: bool = False) -> List[Tuple[str, Tuple[int, ...]]]:
    """Variable (name, shape) list in the order Keras creates them == checkpoint order
    (Results/**/checkpoint-1250: fc_1, fc_value_1, fc_2, fc_value_2, fc_out, value_out).
    For depth n the interleaving is fc_i, fc_value_i per depth, then fc_out, value_out."""
    out = []
    n_out = num_outputs // 2 if free_log_std else num_outputs
    if free_log_std:
        out.append(("log_std", (n_out,)))
    prev = D
    for i, h in enumerate(hiddens, start=1):
        out.append((f"fc_{i}/kernel", (prev, h)))
        out.append((f"fc_{i}/bias", (h,)))
        if not vf_share_layers:
            out.append((f"fc_value_{i}/kernel", (prev, h)))
            out.append((f"fc_value_{i}/bias", (h,)))
        prev = h
    out.append(("fc_out/kernel", (prev, n_out)))
    out.append(("fc_out/bias", (n_out,)))
    out.append(("value_out/kernel", (prev, 1)))
    out.append(("value_out/bias", (1,)))
    return out


def n_params(shapes) -> int:
    return int(sum(int(np.prod(s)) for _, s in shapes))


def unflatten(theta: torch.Tensor, shapes) -> Dict[str, torch.Tensor]:
    out, o = {}, 0
    for name, shp in shapes:
        n = int(np.prod(shp))
        out[name] = theta[o:o + n].reshape(shp)
        o += n
    assert o == theta.numel(), (o, theta.numel())
    return out


def fcnet_init(D: int, num_outputs: int, gen: torch.Generator, hiddens=(64, 64), vf_share_layers=False,
               free_log_std=False, dtype=torch.float32) -> torch.Tensor:
    """Glorot scale 1.0 for hidden layers, 0.01 for fc_out / value_out, zero biases
    (fcnet_glorot_uniform_init.py:53,72,78,105,112)."""
    chunks = []
    for name, shp in fcnet_shapes(D, num_outputs, hiddens, vf_share_layers, free_log_std):
        if len(shp) == 1:
            chunks.append(torch.zeros(shp, dtype=dtype))
        else:
            scale = 0.01 if name.startswith(("fc_out", "value_out")) else 1.0
            chunks.append(glorot_uniform_scaled(shp[0], shp[1], scale, gen, dtype).reshape(-1))
    return torch.cat([c.reshape(-1) for c in chunks])


def _act(name: Optional[str]):
    return {"tanh": torch.tanh, "relu": torch.relu, "linear": (lambda x: x), None: (lambda x: x)}[name]


def fcnet_forward(theta: torch.Tensor, x: torch.Tensor, num_outputs: int, hiddens=(64, 64),
                  activation="tanh", vf_share_layers=False, free_log_std=False):
    """-> (logits[B,num_outputs], value[B]).  fcnet_glorot_uniform_init.py:48-78 (policy branch),
    :95-113 (value branch), :120-125 (forward / value_function)."""
    D = x.shape[-1]
    p = unflatten(theta, fcnet_shapes(D, num_outputs, hiddens, vf_share_layers, free_log_std))
    act = _act(activation)
    h = x
    g = x
    for i in range(1, len(hiddens) + 1):
        h = act(h @ p[f"fc_{i}/kernel"] + p[f"fc_{i}/bias"])
        if not vf_share_layers:
            g = act(g @ p[f"fc_value_{i}/kernel"] + p[f"fc_value_{i}/bias"])
    logits = h @ p["fc_out/kernel"] + p["fc_out/bias"]
    if free_log_std:
        logits = torch.cat([logits, p["log_std"].expand(x.shape[0], -1)], dim=1)
    vin = h if vf_share_layers else g
    value = (vin @ p["value_out/kernel"] + p["value_out/bias"]).reshape(-1)
    return logits, value


# --------------------------------------------------------------------------------------------
# a12 LegCoupling                                   models/coupling_net_glorot_uniform_init.py:11-30
# --------------------------------------------------------------------------------------------

COUPLING_INIT = ((1.0, 1.0), (-1.0, -1.0), (-1.0, -1.0), (1.0, 1.0))  # coupling_net…:20-21


def leg_coupling(logits: torch.Tensor, node_id: torch.Tensor, coupling: torch.Tensor) -> torch.Tensor:
    """logits[B,2A] * pad(coupling[4,2], ones -> [4, 2 + 2A//2])[node_id]   (coupling_net…:28-30).
    The pad width is n_dims = input_shape[-1]//2, so the product broadcasts only when 2 + A == 2A,
    i.e. A == 2 (the per-leg action space) — same restriction as the reference."""
    n_dims = logits.shape[-1] // 2
    ones = torch.ones(coupling.shape[0], n_dims, dtype=logits.dtype)
    c = torch.cat([coupling.to(logits.dtype), ones], dim=1)
    return logits * c[node_id.reshape(-1).long()]


# --------------------------------------------------------------------------------------------
# a11 graph_ops + graph layers                      models/graph_ops.py:3-26, models/gcn.py:7-206
# --------------------------------------------------------------------------------------------


def adj_norm(adj: torch.Tensor) -> torch.Tensor:
    """D^-1 A with D = rowsum(A)            (graph_ops.py:13-21)."""
    d = adj.sum(-1) ** -1.0
    return torch.diag_embed(d) @ adj


def symm_norm(adj: torch.Tensor) -> torch.Tensor:
    """D^-1/2 A D^-1/2                      (graph_ops.py:3-11)."""
    d = adj.sum(-1) ** -0.5
    return torch.diag_embed(d) @ adj @ torch.diag_embed(d)


def segment_softmax(data: torch.Tensor, segment_ids: torch.Tensor, num_segments: int) -> torch.Tensor:
    """exp(data) / segment_sum(exp(data))[ids]   (graph_ops.py:23-26; no max-subtraction, as there)."""
    e = torch.exp(data)
    sums = torch.zeros((num_segments,) + tuple(data.shape[1:]), dtype=data.dtype).index_add_(0, segment_ids, e)
    return e / sums[segment_ids]


def gcn_layer(x, adj, W, b=None, activation="tanh"):
    """act((D^-1 A) X W + b)                (gcn.py:29-37)."""
    y = (adj_norm(adj) @ x) @ W
    if b is not None:
        y = y + b
    return _act(activation)(y)


def _edges(adj):
    """tf.where(adj) -> (batch, sender, receiver), row-major order (gcn.py:60-62)."""
    e = torch.nonzero(adj != 0)
    return e[:, 0], e[:, 1], e[:, 2]


def _segment_mean(data, ids, num):
    """tf.math.unsorted_segment_mean: sum / max(count, 1)."""
    s = torch.zeros((num, data.shape[1]), dtype=data.dtype).index_add_(0, ids, data)
    c = torch.zeros(num, dtype=data.dtype).index_add_(0, ids, torch.ones(ids.shape[0], dtype=data.dtype))
    return s / torch.clamp(c, min=1.0).unsqueeze(1)


def mpnn_layer(x, adj, W_msg, W_upd, b=None, activation="tanh"):
    """act(X W_upd + segment_mean_{s->r}(X_s W_msg) [+ b])     (gcn.py:57-94).
    Edge convention adj[b, s, r] != 0 => message s -> r; receivers without in-edges get 0."""
    B, n = adj.shape[:2]
    bt, snd, rcv = _edges(adj)
    x_snd = x[bt, snd] @ W_msg
    msgs = _segment_mean(x_snd, rcv + bt * n, n * B).reshape(B, n, -1)
    y = x @ W_upd + msgs
    if b is not None:
        y = y + b
    return _act(activation)(y)


def mpnn2_layer(x, adj, W_msg, W_upd, b=None, activation="tanh"):
    """Concat variant: e = [x_s, x_r] W_msg; y = act([x, mean(e)] W_upd [+ b])   (gcn.py:113-150)."""
    B, n = adj.shape[:2]
    bt, snd, rcv = _edges(adj)
    e = torch.cat([x[bt, snd], x[bt, rcv]], dim=-1) @ W_msg
    msgs = _segment_mean(e, rcv + bt * n, n * B).reshape(B, n, -1)
    y = torch.cat([x, msgs], dim=-1) @ W_upd
    if b is not None:
        y = y + b
    return _act(activation)(y)


def gat1_layer(x, adj, W_pre, w_att, b=None, activation="tanh"):
    """Attention variant with self loops     (gcn.py:171-206)."""
    B, n = adj.shape[:2]
    adj = torch.minimum(torch.ones((), dtype=adj.dtype), adj + torch.eye(n, dtype=adj.dtype)[None])
    bt, snd, rcv = _edges(adj)
    x = x @ W_pre
    pre = torch.cat([x[bt, snd], x[bt, rcv]], dim=-1)
    att = torch.nn.functional.leaky_relu(pre @ w_att, negative_slope=0.2)  # tf.nn.leaky_relu alpha=0.2
    att = segment_softmax(att, rcv + bt * n, n * B)
    A = torch.zeros_like(adj)
    A[bt, snd, rcv] = att[:, 0]
    y = A @ x
    if b is not None:
        y = y + b
    return _act(activation)(y)


# --------------------------------------------------------------------------------------------
# a10 GraphNet + wrapper       models/graph_net.py:10-45, models/shared_graphnet_glorot_uniform_init.py:21-58
# --------------------------------------------------------------------------------------------

GN_LEG_FEATS = 19  # hard-coded in graph_net.py:16,35
GN_ENC_IN = 4      # state[..., -4:]  graph_net.py:33


def graphnet_shapes(num_outputs: int, hiddens=(64, 64)):
    """Keras variable order of GraphNet: enc(kernel,bias), gnn.msg_transform, gnn.node_update,
    out(kernel,bias)   (graph_net.py:14-29; MPNN has use_bias=False, :23)."""
    H0, H1 = hiddens
    return [("state_enc/kernel", (GN_ENC_IN, GN_LEG_FEATS * H0)), ("state_enc/bias", (GN_LEG_FEATS * H0,)),
            ("msg_transform/kernel", (H0, H1)), ("node_update/kernel", (H0, H1)),
            ("linear_out/kernel", (H1, num_outputs)), ("linear_out/bias", (num_outputs,))]


def graphnet_init(num_outputs: int, gen, hiddens=(64, 64), dtype=torch.float32):
    chunks = []
    for name, shp in graphnet_shapes(num_outputs, hiddens):
        if len(shp) == 1:
            chunks.append(torch.zeros(shp, dtype=dtype))
        else:
            scale = 0.01 if name.startswith("linear_out") else 1.0
            chunks.append(glorot_uniform_scaled(shp[0], shp[1], scale, gen, dtype).reshape(-1))
    return torch.cat(chunks)


def graphnet_enc_leg_features(p, state, H0=64):
    """w = tanh(Dense(state[..., -4:])) -> [B,n,19,H0];  x = tanh(state[..., :-4] @ w)   (graph_net.py:32-37)."""
    w = torch.tanh(state[..., -GN_ENC_IN:] @ p["state_enc/kernel"] + p["state_enc/bias"])
    B, n = state.shape[:2]
    w = w.reshape(B, n, GN_LEG_FEATS, H0)
    leg = state[..., :-GN_ENC_IN].unsqueeze(-2)
    return torch.tanh((leg @ w).squeeze(-2))


def graphnet_forward_one(theta, node_idx, state, adj, num_outputs: int, hiddens=(64, 64)):
    """GraphNet.call (graph_net.py:39-45) -> [B, num_outputs]."""
    p = unflatten(theta, graphnet_shapes(num_outputs, hiddens))
    x = graphnet_enc_leg_features(p, state, hiddens[0])
    x = mpnn_layer(x, adj, p["msg_transform/kernel"], p["node_update/kernel"], None, "tanh")
    idx = node_idx.reshape(-1).long()
    x = x[torch.arange(x.shape[0]), idx]
    return x @ p["linear_out/kernel"] + p["linear_out/bias"]


def graphnet_forward(theta, node_idx, state, adj, num_outputs: int, hiddens=(64, 64)):
    """Wrapper model: separate actor GraphNet(num_outputs) and critic GraphNet(1); theta = [actor|critic]
    (shared_graphnet…:32-33,52-58) -> (logits[B,num_outputs], value[B])."""
    na = n_params(graphnet_shapes(num_outputs, hiddens))
    logits = graphnet_forward_one(theta[:na], node_idx, state, adj, num_outputs, hiddens)
    value = graphnet_forward_one(theta[na:], node_idx, state, adj, 1, hiddens).reshape(-1)
    return logits, value


def graphnet_wrapper_init(num_outputs, gen, hiddens=(64, 64), dtype=torch.float32):
    return torch.cat([graphnet_init(num_outputs, gen, hiddens, dtype), graphnet_init(1, gen, hiddens, dtype)])


def ring_adjacency(dtype=torch.float32):
    """FL-HL-HR-FR-FL both directions, no self loops
    (quantruped_GraphDecentralizedController_environments.py:167-190); node order FL,HL,HR,FR."""
    a = torch.zeros(4, 4, dtype=dtype)
    for s, r in [(0, 1), (1, 2), (2, 3), (3, 0), (1, 0), (2, 1), (3, 2), (0, 3)]:
        a[s, r] = 1.0
    return a


# --------------------------------------------------------------------------------------------
# a3  obs index tables          simulation_envs/quantruped_v3.py:68-102,282-317
# --------------------------------------------------------------------------------------------

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


def get_obs_indices(prefixes, use_target_velocity=False):
    """Prefix-major order (quantruped_v3.py:282-300); TVel appends 'body_target_x_vel' as index 43
    (quantruped_v3.py:394-400), which the 'body' prefix then matches."""
    fields = OBS_FIELDS + (["body_target_x_vel"] if use_target_velocity else [])
    if prefixes is None:
        return list(range(len(fields)))
    out = []
    for p in prefixes:
        out.extend(i for i, f in enumerate(fields) if f.startswith(p))
    return out


def get_action_indices(prefixes):
    out = []
    for p in prefixes:
        out.extend(i for i, f in enumerate(ACTION_FIELDS) if f.startswith(p))
    return out


def get_contact_force_indices(prefixes=None, weights=None):
    """simulation_envs/quantruped_v3.py:319-341."""
    fields = ['body_floor', 'body', 'fl_hip', 'fl_leg', 'fl_foot', 'hl_hip', 'hl_leg', 'hl_foot',
              'hr_hip', 'hr_leg', 'hr_foot', 'fr_hip', 'fr_leg', 'fr_foot']      # quantruped_v3.py:105-112
    if prefixes is None:
        return np.arange(len(fields)), np.ones([len(fields), 1])
    if weights is None:
        weights = np.ones(len(prefixes))
    idx, w = [], []
    for prefix, weight in zip(prefixes, weights):
        hit = list(np.where([f.startswith(prefix) for f in fields])[0])
        idx.extend(hit)
        w.extend([[weight]] * len(hit))
    return idx, w


def distribute_rewards(fw_reward, action_dict, cfrc_ext, contact_force_indices, agent_names, ctrl_cost_weight,
                       contact_cost_weight, mode="per_leg"):
    """One env step of the adaptor's reward split, line by line (quantruped_adaptor_multi_environment.py:151-203;
    mode 'global_costs': quantruped_fourDecentralizedController_GlobalCosts_environments.py:69-83)  -> {agent: reward}."""
    contact_forces = np.clip(cfrc_ext, -1., 1.)
    contact_costs = contact_cost_weight * np.square(contact_forces)
    contact_cost = {}
    for a in agent_names:
        idx, weights = contact_force_indices[a]
        contact_cost[a] = np.sum(np.multiply(contact_costs[idx], weights))
    rew = {}
    n = len(agent_names)
    if mode == "global":
        contact_costs_sum = contact_cost_weight * np.sum(np.square(contact_forces))
        ctrl_costs_sum = 0.
        for a in agent_names:
            ctrl_costs_sum += np.sum(np.square(action_dict[a]))
        for a in agent_names:
            rew[a] = (fw_reward - ctrl_cost_weight * ctrl_costs_sum - contact_costs_sum) / n
    elif mode == "global_costs":
        s = 0
        for a in agent_names:
            s += np.sum(np.square(action_dict[a]))
        for a in agent_names:
            rew[a] = (fw_reward / n) - (ctrl_cost_weight * 0.25 * s) - contact_cost[a]
    else:
        for a in agent_names:
            if mode == "per_leg_norm":
                rew[a] = fw_reward - n * (ctrl_cost_weight * np.sum(np.square(action_dict[a])) + contact_cost[a])
            else:
                rew[a] = fw_reward / n - ctrl_cost_weight * np.sum(np.square(action_dict[a])) - contact_cost[a]
    return rew


def concatenate_actions(action_dict, action_indices):
    """quantruped_adaptor_multi_environment.py:205-212."""
    actions = np.empty(8,)
    for k in action_dict:
        actions[action_indices[k]] = action_dict[k]
    return actions


LEG_ANGLES = {"agent_FL": 45., "agent_HL": 135., "agent_HR": -135., "agent_FR": -45.}   # …GraphDecentralized…:138-143


def leg_encoding(angle):
    """quantruped_GraphDecentralizedController_environments.py:145-147  -> [sin, cos] of the angle in degrees."""
    rad = np.deg2rad(angle)
    return np.stack((np.sin(rad), np.cos(rad)))


def quaternion_multiply(quat1, quat2):
    """quantruped_GraphDecentralizedController_environments.py:149-156 (component order as written there)."""
    x1, y1, z1, w1 = quat1
    x2, y2, z2, w2 = quat2
    return np.array([x1 * w2 + y1 * z2 - z1 * y2 + w1 * x2,
                     -x1 * z2 + y1 * w2 + z1 * x2 + w1 * y2,
                     x1 * y2 - y1 * x2 + z1 * w2 + w1 * z2,
                     -x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2])


def leg_encoding_ego(angle, full_obs):
    """:158-161 — body orientation obs[1:5] times the leg's yaw quaternion (half angle)."""
    quat_z, quat_w = leg_encoding(angle / 2.)
    return quaternion_multiply(full_obs[1:5], [0., 0., quat_z, quat_w])


def graph_distribute_observations(obs_full, normalize, obs_indices, leg_angles=None):
    """Node-feature matrix of the shared-graph env for ONE env step
    (quantruped_GraphDecentralizedController_environments.py:215-245): per agent, the normalised 43-dim observation
    gathered by the agent's index list, followed by the 4 ego leg-encoding numbers computed from the RAW observation.
    `normalize` is the env-side filter call (`_normalize_observation`, quantruped_adaptor_multi_environment.py:83-85).
    Returns [n_agents, len(idx) + 4] in obs_full's dtype; every agent receives (its index, this matrix, adj)."""
    leg_angles = LEG_ANGLES if leg_angles is None else leg_angles
    normed = normalize(obs_full)
    rows = [np.concatenate((normed[obs_indices[a]], leg_encoding_ego(leg_angles[a], obs_full))) for a in obs_indices]
    return np.stack(rows).astype(obs_full.dtype)


# --------------------------------------------------------------------------------------------
# a4  MeanStdFilter / RunningStat        ray.rllib.utils.filter (1.0.1), used at
#     simulation_envs/observation_filter.py:8-12 and via observation_filter="MeanStdFilter"
# --------------------------------------------------------------------------------------------


class RunningStat:
    def __init__(self, shape):
        self._n = 0
        self._M = np.zeros(shape, dtype=np.float64)
        self._S = np.zeros(shape, dtype=np.float64)

    def push(self, x):
        x = np.asarray(x, dtype=np.float64)
        assert x.shape == self._M.shape
        n1 = self._n
        self._n += 1
        if self._n == 1:
            self._M[...] = x
        else:
            delta = x - self._M
            self._M[...] += delta / self._n
            self._S[...] += delta * delta * n1 / self._n

    def update(self, other: "RunningStat"):
        n1, n2 = self._n, other._n
        n = n1 + n2
        if n == 0:
            return
        delta = self._M - other._M
        M = (n1 * self._M + n2 * other._M) / n
        S = self._S + other._S + delta * delta * n1 * n2 / n
        self._n, self._M, self._S = n, M, S

    @property
    def n(self):
        return self._n

    @property
    def mean(self):
        return self._M

    @property
    def var(self):
        return self._S / (self._n - 1) if self._n > 1 else np.square(self._M)

    @property
    def std(self):
        return np.sqrt(self.var)


class MeanStdFilter:
    """y = clip((x - mean) / (std + 1e-8)).  A call with one extra leading axis pushes EVERY row
    first and then normalises all rows with the final statistics."""

    def __init__(self, shape, demean=True, destd=True, clip=10.0):
        self.shape, self.demean, self.destd, self.clip = shape, demean, destd, clip
        self.rs = RunningStat(shape)
        self.buffer = RunningStat(shape)

    def __call__(self, x, update=True):
        x = np.asarray(x, dtype=np.float64)
        if update:
            if x.ndim == len(self.rs._M.shape) + 1:
                for i in range(x.shape[0]):
                    self.rs.push(x[i])
                    self.buffer.push(x[i])
            else:
                self.rs.push(x)
                self.buffer.push(x)
        if self.demean:
            x = x - self.rs.mean
        if self.destd:
            x = x / (self.rs.std + 1e-8)
        if self.clip:
            x = np.clip(x, -self.clip, self.clip)
        return x


def batch_stat(x: np.ndarray) -> RunningStat:
    """Vectorised (two-pass) statistics of rows of x — what a parallel device reduction computes."""
    rs = RunningStat(x.shape[1:])
    rs._n = int(x.shape[0])
    if rs._n:
        rs._M = x.mean(axis=0, dtype=np.float64)
        rs._S = ((x - rs._M) ** 2).sum(axis=0, dtype=np.float64)
    return rs


# --------------------------------------------------------------------------------------------
# a5  DiagGaussian                        ray.rllib.models.tf.tf_action_dist.DiagGaussian (1.0.1)
# --------------------------------------------------------------------------------------------


def dg_split(logits):
    A = logits.shape[-1] // 2
    return logits[..., :A], logits[..., A:]


def dg_sample(logits, eps):
    mean, log_std = dg_split(logits)
    return mean + torch.exp(log_std) * eps


def dg_logp(logits, a):
    mean, log_std = dg_split(logits)
    A = mean.shape[-1]
    return (-0.5 * (((a - mean) / torch.exp(log_std)) ** 2).sum(-1) - 0.5 * LOG_2PI * A - log_std.sum(-1))


def dg_kl(logits_p, logits_q):
    """KL(p || q)."""
    mp, lp = dg_split(logits_p)
    mq, lq = dg_split(logits_q)
    return (lq - lp + (torch.exp(lp) ** 2 + (mp - mq) ** 2) / (2.0 * torch.exp(lq) ** 2) - 0.5).sum(-1)


def dg_entropy(logits):
    _, log_std = dg_split(logits)
    return (log_std + 0.5 * math.log(2.0 * math.pi * math.e)).sum(-1)


# --------------------------------------------------------------------------------------------
# a6  GAE            ray.rllib.evaluation.postprocessing.compute_advantages (1.0.1), float64
# --------------------------------------------------------------------------------------------


def discount(x: np.ndarray, gamma: float) -> np.ndarray:
    import scipy.signal
    return scipy.signal.lfilter([1], [1, float(-gamma)], x[::-1], axis=0)[::-1]


def compute_advantages_fragment(rewards, vf_preds, last_r, gamma=0.99, lambda_=0.95):
    """One fragment of one (env, agent).  -> (advantages f32, value_targets f32)."""
    vpred_t = np.concatenate([np.asarray(vf_preds, dtype=np.float32), np.array([last_r])]).astype(np.float64)
    delta_t = np.asarray(rewards, dtype=np.float64) + gamma * vpred_t[1:] - vpred_t[:-1]
    adv = discount(delta_t, gamma * lambda_)
    vtarg = (adv + np.asarray(vf_preds, dtype=np.float32)).astype(np.float32)
    return adv.astype(np.float32), vtarg


def gae_columns(rewards, values, dones, v_boot, gamma=0.99, lambda_=0.95):
    """rewards, values [T, C]; dones [T, C] (1 = episode ended AT this step); v_boot [C] = V(s_T).
    Splits every column into fragments at dones and applies the reference procedure to each."""
    T, C = rewards.shape
    adv = np.zeros((T, C), np.float32)
    vt = np.zeros((T, C), np.float32)
    for c in range(C):
        start = 0
        for t in range(T):
            end_here = bool(dones[t, c]) or t == T - 1
            if end_here:
                last_r = 0.0 if dones[t, c] else float(v_boot[c])
                a, v = compute_advantages_fragment(rewards[start:t + 1, c], values[start:t + 1, c], last_r, gamma, lambda_)
                adv[start:t + 1, c], vt[start:t + 1, c] = a, v
                start = t + 1
    return adv, vt


def gae_recurrence(rewards, values, dones, v_boot, gamma=0.99, lambda_=0.95):
    """Batched equivalent (what the device kernel runs): A_t = d_t + g*l*(1-done_t)*A_{t+1}."""
    T, C = rewards.shape
    r = rewards.astype(np.float64)
    v = values.astype(np.float32).astype(np.float64)
    nd = 1.0 - dones.astype(np.float64)
    adv = np.zeros((T, C), np.float64)
    nxt_v = v_boot.astype(np.float32).astype(np.float64)
    nxt_a = np.zeros(C, np.float64)
    for t in range(T - 1, -1, -1):
        delta = r[t] + gamma * nd[t] * nxt_v - v[t]
        nxt_a = delta + gamma * lambda_ * nd[t] * nxt_a
        adv[t] = nxt_a
        nxt_v = v[t]
    return adv.astype(np.float32), (adv + v).astype(np.float32)


# --------------------------------------------------------------------------------------------
# a7  StandardizeFields(["advantages"])        ray.rllib.execution.rollout_ops (1.0.1)
# --------------------------------------------------------------------------------------------


def standardized(a: np.ndarray) -> np.ndarray:
    return (a - a.mean()) / max(1e-4, a.std())


# --------------------------------------------------------------------------------------------
# a8  PPOLoss                               ray.rllib.agents.ppo.ppo_tf_policy.PPOLoss (1.0.1)
# --------------------------------------------------------------------------------------------


@dataclass
class PPOConfig:
    """Resolved values of every published run (Results/**/params.json)."""
    gamma: float = 0.99
    lambda_: float = 0.95
    clip_param: float = 0.2
    vf_clip_param: float = 10.0
    vf_loss_coeff: float = 0.5
    entropy_coeff: float = 0.0
    kl_coeff: float = 0.2
    kl_target: float = 0.01
    lr: float = 3e-4
    grad_clip: float = 0.5
    num_sgd_iter: int = 10
    sgd_minibatch_size: int = 128
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8


def ppo_loss_from_outputs(logits, value, actions, old_logits, old_logp, vf_preds, advantages, value_targets,
                          kl_coeff: float, cfg: PPOConfig):
    """-> (total_loss, stats dict of 0-d tensors)."""
    logp = dg_logp(logits, actions)
    ratio = torch.exp(logp - old_logp)
    kl = dg_kl(old_logits, logits)
    ent = dg_entropy(logits)
    surr = torch.minimum(advantages * ratio, advantages * torch.clamp(ratio, 1.0 - cfg.clip_param, 1.0 + cfg.clip_param))
    vf1 = (value - value_targets) ** 2
    vclip = vf_preds + torch.clamp(value - vf_preds, -cfg.vf_clip_param, cfg.vf_clip_param)
    vf2 = (vclip - value_targets) ** 2
    vf = torch.maximum(vf1, vf2)
    total = (-surr + kl_coeff * kl + cfg.vf_loss_coeff * vf - cfg.entropy_coeff * ent).mean()
    yvar = value_targets.var(unbiased=False)
    dvar = (value_targets - value).var(unbiased=False)
    stats = {
        "total_loss": total, "policy_loss": (-surr).mean(), "vf_loss": vf.mean(), "kl": kl.mean(),
        "entropy": ent.mean(),
        "vf_explained_var": torch.maximum(torch.tensor(-1.0, dtype=value.dtype), 1.0 - dvar / yvar),
    }
    return total, stats


# --------------------------------------------------------------------------------------------
# a9  clip_by_global_norm, TF1 Adam, minibatch SGD loop, KL-coefficient update
# --------------------------------------------------------------------------------------------


def clip_by_global_norm(g: torch.Tensor, clip: float):
    """tf.clip_by_global_norm on the concatenated gradient: g * clip / max(||g||, clip)."""
    norm = torch.sqrt((g * g).sum())
    scale = clip * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip, dtype=g.dtype))
    return g * scale, norm


@dataclass
class AdamState:
    m: torch.Tensor
    v: torch.Tensor
    beta1_power: float
    beta2_power: float

    @staticmethod
    def zeros(n, dtype, cfg: PPOConfig):
        return AdamState(torch.zeros(n, dtype=dtype), torch.zeros(n, dtype=dtype), cfg.beta1, cfg.beta2)


def adam_tf1_step(theta, g, st: AdamState, cfg: PPOConfig):
    """tf.compat.v1.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps);
    beta powers are multiplied AFTER the update.  Scalars are rounded to the tensor dtype the way the
    FP32 TF kernel sees them."""
    dt = theta.dtype
    f = (lambda x: float(np.float32(x))) if dt == torch.float32 else float
    b1p, b2p = f(st.beta1_power), f(st.beta2_power)
    lr_t = f(f(cfg.lr) * f(math.sqrt(f(1.0 - b2p))) / f(1.0 - b1p))
    # tensorflow/core/kernels/training_ops.cc ApplyAdam (non-nesterov):
    #   m += (g - m)*(1-b1);  v += (g*g - v)*(1-b2);  var -= (m*alpha)/(sqrt(v)+eps)
    st.m = st.m + (g - st.m) * f(1.0 - f(cfg.beta1))
    st.v = st.v + (g * g - st.v) * f(1.0 - f(cfg.beta2))
    theta = theta - (st.m * lr_t) / (torch.sqrt(st.v) + f(cfg.adam_eps))
    st.beta1_power = f(b1p * f(cfg.beta1))
    st.beta2_power = f(b2p * f(cfg.beta2))
    return theta


def update_kl(kl_coeff: float, sampled_kl: float, kl_target: float) -> float:
    """KLCoeffMixin.update_kl."""
    if sampled_kl > 2.0 * kl_target:
        return kl_coeff * 1.5
    if sampled_kl < 0.5 * kl_target:
        return kl_coeff * 0.5
    return kl_coeff


STAT_KEYS = ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "vf_explained_var")


def sgd_minibatch_step(theta, st: AdamState, forward_fn, batch: Dict[str, torch.Tensor], rows: slice,
                       kl_coeff: float, cfg: PPOConfig):
    """One optimizer step on batch[rows]: loss -> autograd -> clip -> Adam.  -> (theta, stats, grad, gnorm)."""
    th = theta.detach().clone().requires_grad_(True)
    logits, value = forward_fn(th, batch["obs"][rows])
    loss, stats = ppo_loss_from_outputs(
        logits, value, batch["actions"][rows], batch["old_logits"][rows], batch["old_logp"][rows],
        batch["vf_preds"][rows], batch["advantages"][rows], batch["value_targets"][rows], kl_coeff, cfg)
    (g,) = torch.autograd.grad(loss, th)
    gc, gnorm = clip_by_global_norm(g, cfg.grad_clip)
    new_theta = adam_tf1_step(theta.detach(), gc, st, cfg)
    return new_theta, {k: float(v.detach()) for k, v in stats.items()}, g, float(gnorm)


def sgd_loop(theta, st: AdamState, forward_fn, batch: Dict[str, torch.Tensor], perms: np.ndarray,
             kl_coeff: float, cfg: PPOConfig):
    """TrainTFMultiGPU inner loop for ONE policy: perms[E, nb] holds the minibatch visiting order of
    every epoch (np.random.permutation(nb) in the reference); minibatch b = rows [b*MB, (b+1)*MB).
    Returned stats = mean over the minibatches of the LAST epoch."""
    MB = cfg.sgd_minibatch_size
    last = None
    for e in range(perms.shape[0]):
        acc = {k: [] for k in STAT_KEYS}
        for b in perms[e]:
            theta, s, _, _ = sgd_minibatch_step(theta, st, forward_fn, batch, slice(int(b) * MB, (int(b) + 1) * MB), kl_coeff, cfg)
            for k in STAT_KEYS:
                acc[k].append(s[k])
        last = {k: float(np.mean(np.asarray(v, dtype=np.float32))) for k, v in acc.items()}
    return theta, last


# --------------------------------------------------------------------------------------------
# Whole learner iteration for grouped FCNet policies (SURVEY.md §8-d "one learner iteration")
# --------------------------------------------------------------------------------------------


@dataclass
class PolicyState:
    theta: torch.Tensor
    adam: AdamState
    filt: MeanStdFilter
    kl_coeff: float


def fcnet_learner_iteration(pols: List[PolicyState], raw_obs: np.ndarray, boot_obs: np.ndarray,
                            rewards: np.ndarray, dones: np.ndarray, eps: np.ndarray, shuffle: Optional[np.ndarray],
                            perms: np.ndarray, num_outputs: int, cfg: PPOConfig, dtype=torch.float32):
    """raw_obs [P,T,C,D] (already index-gathered per policy), boot_obs [P,C,D] = obs after the last step,
    rewards [P,T,C], dones [T,C] (shared by the policies of one env; C = N * agents-per-policy columns,
    env-major), eps [P,T,C,A], shuffle [P,T*C] row permutation or None, perms [P,E,nb].
    Steps: filter update+normalise -> forward/sample/logp/value -> bootstrap -> GAE -> standardise ->
    E epochs of minibatch SGD -> KL-coefficient update.  Returns per-policy dict of results."""
    P, T, C, D = raw_obs.shape
    out = []
    for p in range(P):
        ps = pols[p]
        obs = ps.filt(raw_obs[p].reshape(T * C, D))                 # pushes all rows, then normalises
        bobs = ps.filt(boot_obs[p], update=False)
        x = torch.from_numpy(obs.astype(np.float32)).to(dtype)
        xb = torch.from_numpy(bobs.astype(np.float32)).to(dtype)
        with torch.no_grad():
            logits, value = fcnet_forward(ps.theta, x, num_outputs)
            act = dg_sample(logits, torch.from_numpy(eps[p].reshape(T * C, -1)).to(dtype))
            logp = dg_logp(logits, act)
            _, vboot = fcnet_forward(ps.theta, xb, num_outputs)
        adv, vt = gae_recurrence(rewards[p], value.numpy().astype(np.float32).reshape(T, C), dones,
                                 vboot.numpy().astype(np.float32), cfg.gamma, cfg.lambda_)
        adv_std = standardized(adv.reshape(-1))
        batch = {"obs": x, "actions": act, "old_logits": logits, "old_logp": logp, "vf_preds": value,
                 "advantages": torch.from_numpy(adv_std).to(dtype),
                 "value_targets": torch.from_numpy(vt.reshape(-1)).to(dtype)}
        if shuffle is not None:
            idx = torch.from_numpy(shuffle[p].astype(np.int64))
            batch = {k: v[idx] for k, v in batch.items()}
        fwd = lambda th, xx: fcnet_forward(th, xx, num_outputs)
        ps.theta, stats = sgd_loop(ps.theta, ps.adam, fwd, batch, perms[p], ps.kl_coeff, cfg)
        ps.kl_coeff = update_kl(ps.kl_coeff, stats["kl"], cfg.kl_target)
        out.append({"stats": stats, "logits": logits, "value": value, "actions": act, "logp": logp,
                    "advantages": adv, "adv_std": adv_std, "value_targets": vt, "obs_norm": obs})
    return out
