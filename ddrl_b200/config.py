"""PPO configuration with the reference's RLlib key names (train_experiment_1_architecture_on_flat.py:96-184;
resolved values of every published run: Results/**/params.json)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict

DEFAULT_MODEL_CONFIG: Dict[str, Any] = {
    "custom_model": "fc_glorot_uniform_init",
    "fcnet_hiddens": [64, 64],
    "fcnet_activation": "tanh",
    "free_log_std": False,
    "no_final_linear": False,
    "vf_share_layers": False,   # top-level vf_share_layers=False overrides the model key in Ray 1.0.x
}

RLLIB_DEFAULTS: Dict[str, Any] = {
    "gamma": 0.99, "lambda": 0.95, "clip_param": 0.2, "vf_clip_param": 10.0, "vf_loss_coeff": 0.5,
    "entropy_coeff": 0.0, "kl_coeff": 0.2, "kl_target": 0.01, "lr": 3e-4, "grad_clip": 0.5,
    "num_sgd_iter": 10, "sgd_minibatch_size": 128, "train_batch_size": 16000, "rollout_fragment_length": 200,
    "observation_filter": "MeanStdFilter", "shuffle_sequences": True, "use_gae": True, "use_critic": True,
    "batch_mode": "truncate_episodes", "clip_actions": True, "normalize_actions": False,
}


@dataclass
class PPOConfig:
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
    beta1: float = 0.9          # tf.compat.v1.train.AdamOptimizer defaults (RLlib TFPolicy.optimizer)
    beta2: float = 0.999
    adam_eps: float = 1e-8
    filter_clip: float = 0.0    # RLlib-side MeanStdFilter has clip=None; env-singleton filter uses 10.0

    @staticmethod
    def from_rllib(cfg: Dict[str, Any]) -> "PPOConfig":
        c = dict(RLLIB_DEFAULTS)
        c.update(cfg or {})
        return PPOConfig(gamma=c["gamma"], lambda_=c["lambda"], clip_param=c["clip_param"],
                         vf_clip_param=c["vf_clip_param"], vf_loss_coeff=c["vf_loss_coeff"],
                         entropy_coeff=c["entropy_coeff"], kl_coeff=c["kl_coeff"], kl_target=c["kl_target"],
                         lr=c["lr"], grad_clip=c["grad_clip"] if c["grad_clip"] is not None else 0.0,
                         num_sgd_iter=c["num_sgd_iter"], sgd_minibatch_size=c["sgd_minibatch_size"])
