#!/usr/bin/env python
"""Golden vectors for the multi-agent adaptor's reward / cost split and action concatenation.

Run HERE (container with /root/reference mounted), never on the GPU box:

    python tests/golden/make_env_glue_golden.py

The reference modules cannot be imported (gym, ray, MuJoCo are absent).  The pure-numpy methods are lifted from the
reference files' ASTs and executed UNMODIFIED on stub objects:
    simulation_envs/quantruped_adaptor_multi_environment.py  QuantrupedMultiPoliciesEnv:
        get_contact_cost_sum, distribute_contact_cost, distribute_global_reward, distribute_per_leg_reward, concatenate_actions
    simulation_envs/quantruped_fourDecentralizedController_GlobalCosts_environments.py  …GlobalCostEnv: distribute_reward
    simulation_envs/quantruped_v3.py  QuAntrupedEnv: get_action_indices, get_contact_force_indices (+ the FIELDS lists)
The stub supplies what the constructors would: `env.sim.data.cfrc_ext`, `env.ctrl_cost_weight`, `env.contact_cost_weight`,
`agent_names`, `action_indices`, `contact_force_indices` (built with the lifted index functions and the constructor
arguments of quantruped_fourDecentralizedController_environments.py:25-36 / …twoDecentralized…:60-69 /
quantruped_centralizedController_environment.py:50-56).  The GlobalCosts `distribute_reward` indexes its dicts by
`policy_names` although they are keyed by agent (a KeyError as shipped); the stub sets policy_names = agent_names,
which is the evident intent.

Writes tests/golden/env_glue.npz with, per architecture tag X in {four, two, one}:
    X/fw [S] f64, X/act [S, Ag, A] f64 (already clipped to [-1, 1]), X/cfrc [S, 14, 6] f64
    X/per_leg, X/per_leg_norm, X/global, X/global_costs   [S, Ag] f64
    X/actions [S, 8] f64;  X/action_idx [Ag, A] i64;  X/contact_w [Ag, 14] f64 dense weight table
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_graph_obs_golden import lift  # noqa: E402

REF = "/root/reference/simulation_envs/"
CTRL_W, CONTACT_W = 0.25, 0.025           # experiment-3 env_config values

ARCHS = {   # tag -> (agent names, action prefixes per agent, contact prefixes / weights per agent)
    "four": (["agent_FL", "agent_HL", "agent_HR", "agent_FR"], [["fl"], ["hl"], ["hr"], ["fr"]],
             [(["body", l], [1. / 4., 1.]) for l in ("fl", "hl", "hr", "fr")]),
    "two": (["agent_LEFT", "agent_RIGHT"], [["fl", "hl"], ["hr", "fr"]],
            [(["body", "fl", "hl"], [1. / 2., 1., 1.]), (["body", "hr", "fr"], [1. / 2., 1., 1.])]),
    "one": (["central_agent"], [None], [(None, None)]),
}


def main():
    Adaptor = lift(REF + "quantruped_adaptor_multi_environment.py", "QuantrupedMultiPoliciesEnv",
                   {"get_contact_cost_sum", "distribute_contact_cost", "distribute_global_reward", "distribute_per_leg_reward",
                    "concatenate_actions"}, set())
    GlobalCost = lift(REF + "quantruped_fourDecentralizedController_GlobalCosts_environments.py",
                      "QuantrupedFullyDecentralizedGlobalCostEnv", {"distribute_reward"}, set())
    sim = lift(REF + "quantruped_v3.py", "QuAntrupedEnv", {"get_action_indices", "get_contact_force_indices"},
               {"ACTION_FIELDS", "CONTACT_FORCE_FIELDS"})()
    out = {}
    S = 48
    for tag, (agents, act_pfx, contact) in ARCHS.items():
        rng = np.random.default_rng(len(agents))
        env = Adaptor()
        env.agent_names = agents
        env.policy_names = agents
        env.action_indices = {a: sim.get_action_indices(p) for a, p in zip(agents, act_pfx)}
        env.contact_force_indices = {a: sim.get_contact_force_indices(p, weights=w) if p is not None
                                     else sim.get_contact_force_indices() for a, (p, w) in zip(agents, contact)}
        env.env = types.SimpleNamespace(ctrl_cost_weight=CTRL_W, contact_cost_weight=CONTACT_W,
                                        sim=types.SimpleNamespace(data=types.SimpleNamespace(cfrc_ext=None)))
        env.distribute_reward_gc = types.MethodType(GlobalCost.distribute_reward, env)
        A = 8 // len(agents)
        fw = rng.standard_normal(S).astype(np.float32).astype(np.float64)          # float32-representable: the device takes float32
        act = np.clip(1.5 * rng.standard_normal((S, len(agents), A)), -1.0, 1.0).astype(np.float32).astype(np.float64)
        cfrc = 2.0 * rng.standard_normal((S, 14, 6))
        cfrc[rng.random((S, 14, 6)) < 0.5] = 0.0
        res = {k: np.zeros((S, len(agents))) for k in ("per_leg", "per_leg_norm", "global", "global_costs")}
        actions = np.zeros((S, 8))
        for s in range(S):
            env.env.sim.data.cfrc_ext = cfrc[s]
            ad = {a: act[s, i] for i, a in enumerate(agents)}
            info = {"reward_forward": fw[s]}
            env.normalize_rewards = False
            r = env.distribute_per_leg_reward(None, info, ad); res["per_leg"][s] = [r[a] for a in agents]
            env.normalize_rewards = True
            r = env.distribute_per_leg_reward(None, info, ad); res["per_leg_norm"][s] = [r[a] for a in agents]
            r = env.distribute_global_reward(None, info, ad); res["global"][s] = [r[a] for a in agents]
            r = env.distribute_reward_gc(None, info, ad); res["global_costs"][s] = [r[a] for a in agents]
            actions[s] = env.concatenate_actions(ad)
        dense = np.zeros((len(agents), 14))
        for i, a in enumerate(agents):
            idx, w = env.contact_force_indices[a]
            for j, wt in zip(np.asarray(idx).tolist(), np.asarray(w).reshape(-1).tolist()):
                dense[i, j] += wt
        out.update({f"{tag}/fw": fw, f"{tag}/act": act, f"{tag}/cfrc": cfrc, f"{tag}/actions": actions,
                    f"{tag}/action_idx": np.asarray([env.action_indices[a] for a in agents], dtype=np.int64),
                    f"{tag}/contact_w": dense, **{f"{tag}/{k}": v for k, v in res.items()}})
    np.savez_compressed(os.path.join(HERE, "env_glue.npz"), **out)
    print("wrote env_glue.npz", sorted(out)[:4], "...")


if __name__ == "__main__":
    main()
