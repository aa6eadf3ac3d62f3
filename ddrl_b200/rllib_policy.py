"""RLlib-facing learner surface: the calls RLlib 1.0.1 makes on a PPO policy, served by the fused
B200 learner for ALL policies of an architecture at once.

RLlib's execution plan (`ray/rllib/agents/ppo/ppo.py` `execution_plan`, driven by the reference through
`tune.run("PPO", config=...)`, train_experiment_1_architecture_on_flat.py:201-211) does, per training iteration:

    rollout workers: obs -> MeanStdFilter -> Policy.compute_actions -> SampleBatch rows
                     Policy.postprocess_trajectory (postprocess_ppo_gae) per episode fragment
    driver:          ConcatBatches -> StandardizeFields(["advantages"]) -> TrainTFMultiGPU / TrainOneStep:
                     for every policy id in the MultiAgentBatch: Policy.learn_on_batch(batch)   (1250 session.run each)
                     UpdateKL

`PPOPolicyGroup` mirrors that surface with the reference's column names (`SampleBatch.OBS` = "obs", "actions",
"action_dist_inputs", "action_logp", "vf_preds", "advantages", "value_targets", "rewards", "dones", "new_obs"):
`compute_actions`, `postprocess_fragments` (batched `postprocess_ppo_gae`), `learn_on_batch` (a MultiAgentBatch-like
`{policy_id: {column: array}}` in, `{policy_id: {"learner_stats": {...}}}` out — the shape `RolloutWorker.learn_on_batch`
returns), `get_weights` / `set_weights` with the TF variable names of the reference checkpoints, and
`update_environment_after_epoch` for the `on_train_result` curriculum callback
(train_experiment_1_architecture_on_flat.py:171-178).  With ray installed the train op of a custom execution plan calls
`learn_on_batch` on the concatenated MultiAgentBatch; ray is absent in this image, so that glue is shown in
INTEGRATION.md §4.1, not shipped.

Everything that computes runs on the GPU through `learner.FCNetLearner` / the C ABI; this module only validates and
stacks host arrays, draws the permutations RLlib would draw (numpy RandomState) and copies to / from the device."""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import policies as _policies
from .config import PPOConfig

OBS, NEXT_OBS, ACTIONS, REWARDS, DONES = "obs", "new_obs", "actions", "rewards", "dones"
ACTION_DIST_INPUTS, ACTION_LOGP, VF_PREDS = "action_dist_inputs", "action_logp", "vf_preds"
ADVANTAGES, VALUE_TARGETS = "advantages", "value_targets"
TRAIN_COLUMNS = (OBS, ACTIONS, ACTION_DIST_INPUTS, ACTION_LOGP, VF_PREDS, ADVANTAGES, VALUE_TARGETS)


class BatchError(ValueError):
    pass


def stack_policy_batches(policy_batches: Mapping[str, Mapping[str, Any]], policy_names: Sequence[str], D: int, A: int,
                         columns: Sequence[str] = TRAIN_COLUMNS) -> Dict[str, np.ndarray]:
    """{policy_id: {column: [R, ...]}} -> {column: float32 [P, R, ...]} in `policy_names` order.

    The fused learner trains the P policies of an architecture in ONE launch, so their batches must have the same
    number of rows — which the reference's environments guarantee (every agent acts at every env step,
    quantruped_adaptor_multi_environment.py:214-246).  Missing policies / columns, ragged row counts, wrong widths and
    non-finite values are errors (RLlib would feed them to TF and train on them)."""
    missing = [p for p in policy_names if p not in policy_batches]
    if missing:
        raise BatchError(f"policy_batches lacks {missing}; policies_to_train = {list(policy_names)}")
    extra = [p for p in policy_batches if p not in policy_names]
    if extra:
        raise BatchError(f"unknown policy ids {extra}; this group trains {list(policy_names)}")
    width = {OBS: D, NEXT_OBS: D, ACTIONS: A, ACTION_DIST_INPUTS: 2 * A}
    out: Dict[str, np.ndarray] = {}
    rows: Optional[int] = None
    for col in columns:
        per_policy = []
        for p in policy_names:
            if col not in policy_batches[p]:
                raise BatchError(f"batch of {p!r} lacks column {col!r}")
            a = np.asarray(policy_batches[p][col])
            want_nd = 2 if col in width else 1
            if a.ndim != want_nd or (col in width and a.shape[1] != width[col]):
                raise BatchError(f"{p}/{col}: expected shape [R{', %d' % width[col] if col in width else ''}], got {a.shape}")
            if rows is None:
                rows = a.shape[0]
            if a.shape[0] != rows:
                raise BatchError(f"{p}/{col} has {a.shape[0]} rows, other columns have {rows}: the grouped learner needs "
                                 "equally long batches for all policies")
            a = a.astype(np.float32, copy=False)
            if not np.isfinite(a).all():
                raise BatchError(f"{p}/{col} contains non-finite values")
            per_policy.append(a)
        out[col] = np.ascontiguousarray(np.stack(per_policy))
    if not rows:
        raise BatchError("empty train batch")
    return out


def draw_minibatch_order(rng: np.random.RandomState, P: int, R: int, num_sgd_iter: int, minibatches: int,
                         shuffle_sequences: bool = True) -> Tuple[Optional[np.ndarray], np.ndarray]:
    """The random draws of RLlib's multi-GPU optimizer, per policy in policy order: one `SampleBatch.shuffle()`
    (`np.random.permutation(R)`, only with `shuffle_sequences`) and one `np.random.permutation(num_batches)` per SGD
    epoch (`TrainTFMultiGPU`, ray/rllib/execution/train_ops.py).  -> shuffle [P, R] int32 | None, perms [P, E, nb] int32."""
    shuffle = None
    if shuffle_sequences:
        shuffle = np.stack([rng.permutation(R) for _ in range(P)]).astype(np.int32)
    perms = np.stack([np.stack([rng.permutation(minibatches) for _ in range(num_sgd_iter)]) for _ in range(P)])
    return shuffle, perms.astype(np.int32)


def usable_rows(R: int, sgd_minibatch_size: int) -> Tuple[int, int]:
    """RLlib's multi-GPU loader truncates the batch to a whole number of minibatches: (rows used, minibatches)."""
    mb = min(int(sgd_minibatch_size), R)
    nb = R // mb
    return nb * mb, nb


def curriculum_smoothness(timesteps_total: float, initial: float, target: float, last_timestep: float, u: float) -> float:
    """Height-field smoothness after a training iteration (`update_environment_after_epoch`,
    quantruped_adaptor_multi_environment.py:97-122), `u` = the `np.random.rand()` draw: inside the curriculum interval a
    random point between flat and the linearly decreasing bound, afterwards anywhere between target and flat."""
    if last_timestep > timesteps_total:
        return initial - u * (initial - target) * (timesteps_total / last_timestep)
    return target + u * (initial - target)


class PPOPolicyGroup:
    """All policies of one `--policy_scope` behind one fused learner; see the module docstring."""

    def __init__(self, policy_scope: str, config: Optional[Dict[str, Any]] = None, device="cuda", seed: Optional[int] = None,
                 use_target_velocity: bool = False, **learner_kw):
        import torch
        from .learner import FCNetLearner
        env = _policies.ARCHITECTURES[policy_scope]
        if getattr(env, "model", "fc_glorot_uniform_init") != "fc_glorot_uniform_init":
            raise NotImplementedError(f"{policy_scope}: PPOPolicyGroup serves the FCNet architectures; the shared GraphNet "
                                      "policy is driven through learner.GraphNetLearner")
        self.env = env
        self.policy_names: List[str] = list(env.policy_names)
        self.config = dict(config or {})
        self.ppo = PPOConfig.from_rllib(self.config)
        self.D, self.A = env.obs_dim(use_target_velocity), env.act_dim()
        self.device = torch.device(device)
        self.rng = np.random.RandomState(self.config.get("seed", seed))
        self.learner = FCNetLearner(len(self.policy_names), self.D, self.A, self.ppo, self.device, **learner_kw)
        self.shuffle_sequences = bool(self.config.get("shuffle_sequences", True))
        self.num_steps_trained = 0
        self.current_smoothness: Optional[float] = None

    # ---- sampler side -------------------------------------------------------------------------------------------------
    def compute_actions(self, obs_by_policy: Mapping[str, np.ndarray], explore: bool = True, update_filter: bool = False):
        """`Policy.compute_actions` for every policy at once: {pid: raw obs [B, D]} ->
        {pid: (actions [B, A], [], {"action_dist_inputs", "action_logp", "vf_preds"})}.  With `explore` actions are
        DiagGaussian samples (noise from this group's RandomState), otherwise the distribution mean (logp of the mean)."""
        import torch
        obs = stack_policy_batches({p: {OBS: obs_by_policy[p]} for p in obs_by_policy}, self.policy_names, self.D, self.A, (OBS,))[OBS]
        P, B = obs.shape[:2]
        noise = self.rng.standard_normal((P, B, self.A)).astype(np.float32) if explore else np.zeros((P, B, self.A), np.float32)
        out = self.learner.compute_actions(torch.from_numpy(obs).to(self.device), torch.from_numpy(noise).to(self.device),
                                           update_filter=update_filter)
        host = {k: v.cpu().numpy() for k, v in out.items() if v is not None}
        res = {}
        for i, p in enumerate(self.policy_names):
            info = {ACTION_DIST_INPUTS: host["logits"][i], ACTION_LOGP: host["logp"][i], VF_PREDS: host["value"][i]}
            res[p] = (host["action"][i], [], info)
        return res

    def postprocess_fragments(self, rewards: np.ndarray, vf_preds: np.ndarray, dones: np.ndarray, last_values: np.ndarray):
        """Batched `postprocess_ppo_gae` (RLlib evaluation/postprocessing.py): rewards / vf_preds [P, T, C] for C
        fragment columns of T steps, dones [T, C] (episode ended AT that step), last_values [P, C] = V(new_obs[-1]) of
        each column (ignored where the column's last step is done) -> (advantages, value_targets) [P, T, C]."""
        import torch
        from . import kernels as K
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        adv, vtarg, _ = K.gae(to(rewards, np.float32), to(vf_preds, np.float32), to(dones, np.uint8), to(last_values, np.float32),
                              1, self.ppo.gamma, self.ppo.lambda_)
        return adv.cpu().numpy(), vtarg.cpu().numpy()

    # ---- learner side -------------------------------------------------------------------------------------------------
    def learn_on_batch(self, policy_batches: Mapping[str, Mapping[str, Any]], standardize: bool = True):
        """`RolloutWorker.learn_on_batch(MultiAgentBatch)`: one PPO update of every policy from postprocessed columns."""
        import torch
        cols = stack_policy_batches(policy_batches, self.policy_names, self.D, self.A)
        R = cols[OBS].shape[1]
        used, nb = usable_rows(R, self.ppo.sgd_minibatch_size)
        P = len(self.policy_names)
        shuffle, perms = draw_minibatch_order(self.rng, P, R, self.ppo.num_sgd_iter, nb, self.shuffle_sequences)
        dev = lambda a: torch.from_numpy(a).to(self.device)
        d = {k: dev(v) for k, v in cols.items()}
        if standardize:        # StandardizeFields sees the WHOLE train batch, before the loader drops the ragged tail
            self.learner.standardize_advantages(d[ADVANTAGES])
        shuf = None if shuffle is None else dev(shuffle)
        if used != R:          # the multi-GPU loader keeps whole minibatches: the first `used` rows of the shuffled batch
            keep = (shuf[:, :used] if shuf is not None else torch.arange(used, device=self.device).expand(P, used)).long()
            d = {k: torch.gather(v, 1, keep.reshape(P, used, *([1] * (v.dim() - 2))).expand(P, used, *v.shape[2:])).contiguous()
                 for k, v in d.items()}
            shuf = None
        stats = self.learner.learn_on_batch(d[OBS], d[ACTIONS], d[ACTION_DIST_INPUTS], d[ACTION_LOGP], d[VF_PREDS],
                                            d[ADVANTAGES], d[VALUE_TARGETS], dev(perms), shuf, standardize=False)
        self.num_steps_trained += used
        return {p: {"learner_stats": stats[i]} for i, p in enumerate(self.policy_names)}

    # ---- weights (TF variable names of the reference checkpoints) -----------------------------------------------------
    def get_weights(self) -> Dict[str, Dict[str, np.ndarray]]:
        from .checkpoint import theta_to_variables
        th = self.learner.theta.cpu().numpy()
        return {p: theta_to_variables(p, th[i], self.D, self.A) for i, p in enumerate(self.policy_names)}

    def set_weights(self, weights: Mapping[str, Mapping[str, np.ndarray]]) -> None:
        import torch
        from .checkpoint import variables_to_theta
        th = np.stack([variables_to_theta(p, weights[p], self.D, self.A) for p in self.policy_names]).astype(np.float32)
        self.learner.theta.copy_(torch.from_numpy(th).to(self.device))

    # ---- curriculum callback ------------------------------------------------------------------------------------------
    def update_environment_after_epoch(self, timesteps_total: float) -> Optional[float]:
        """What `on_train_result` makes every env do (train_experiment_1_architecture_on_flat.py:171-178): returns the new
        height-field smoothness for the simulator (None when `curriculum_learning` is off)."""
        ec = self.config.get("env_config", {})
        if not ec.get("curriculum_learning", False):
            return None
        lo, hi = ec["range_smoothness"]
        self.current_smoothness = curriculum_smoothness(timesteps_total, lo, hi, ec["range_last_timestep"], self.rng.rand())
        return self.current_smoothness
