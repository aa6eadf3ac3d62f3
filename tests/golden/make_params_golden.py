#!/usr/bin/env python
"""Distinct hyper-parameter values of ALL published runs (Results/**/params.json, 120 files) for the keys the learner path
reads — pins `ddrl_b200.config.RLLIB_DEFAULTS` / `PPOConfig.from_rllib` / `DEFAULT_MODEL_CONFIG`.

    python tests/golden/make_params_golden.py        (HERE only: needs /root/reference)

Writes tests/golden/published_params.json: {"n_runs": N, "top": {key: [distinct values]}, "model": {...}, "env_config": {...}}."""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
TOP = ["gamma", "lambda", "clip_param", "vf_clip_param", "vf_loss_coeff", "entropy_coeff", "kl_coeff", "kl_target", "lr",
       "grad_clip", "num_sgd_iter", "sgd_minibatch_size", "train_batch_size", "rollout_fragment_length", "observation_filter",
       "shuffle_sequences", "use_gae", "use_critic", "batch_mode", "clip_actions", "normalize_actions", "vf_share_layers",
       "num_workers", "num_envs_per_worker", "framework", "lr_schedule", "entropy_coeff_schedule"]
MODEL = ["custom_model", "fcnet_hiddens", "fcnet_activation", "free_log_std", "no_final_linear", "vf_share_layers"]


def distinct(values):
    out = []
    for v in values:
        if v not in out:
            out.append(v)
    return sorted(out, key=lambda v: json.dumps(v))


def main():
    files = sorted(glob.glob("/root/reference/Results/**/params.json", recursive=True))
    runs = [json.load(open(f)) for f in files]
    env_keys = sorted({k for r in runs for k in r.get("env_config", {})})
    out = {"n_runs": len(runs),
           "top": {k: distinct(r.get(k, "<absent>") for r in runs) for k in TOP},
           "model": {k: distinct(r["model"].get(k, "<absent>") for r in runs) for k in MODEL},
           "env_config": {k: distinct(r.get("env_config", {}).get(k, "<absent>") for r in runs) for k in env_keys}}
    json.dump(out, open(os.path.join(HERE, "published_params.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out)[:1500])


if __name__ == "__main__":
    main()
