#!/usr/bin/env python
"""Golden vectors of one whole multi-agent env STEP as the reference wires it: the real
constructors and `step()` of every `--policy_scope` class executed unmodified, with only the MuJoCo simulation replaced.

    python tests/golden/make_env_step_golden.py        (HERE only: needs /root/reference)

Lifted per module like tests/golden/make_arch_golden.py, but the root class keeps its own `__init__` / `step` /
`distribute_*` / `concatenate_actions`; `create_env` returns a stub simulator (the index helpers of `QuAntrupedEnv` + a
`step()` that records the action it is given and returns a prepared observation / forward reward / contact forces), and
`MeanStdFilterSingleton` hands out the oracle's MeanStdFilter restatement (RLlib is not installed).
Three env_configs per scope: {} (per-leg reward), {"norm_reward": True}, {"global_reward": True}
(train_experiment_1_architecture_on_flat.py:25-35,155-156).  Each env takes 5 steps (the filter updates on every call).
Writes tests/golden/env_step.npz:
    obs_full [5, 43], fw [5], cfrc [5, 14, 6], actions/<scope> [5, Ag, A]
    <scope>/<cfg>/reward_fn  name of the function `distribute_reward` is bound to, <scope>/<cfg>/normalize_rewards
    <scope>/<cfg>/sim_action [5, 8]   what reached the simulator
    <scope>/<cfg>/rew [5, Ag]         rewards per agent (agent_names order)
    <scope>/<cfg>/obs [5, Ag, D]      observations per agent (graph scopes: the node matrix [5, Ag, 4, F])"""
import os
import random
import sys
import types
from collections.abc import Iterable

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_arch_golden as MA  # noqa: E402
from oracle import ddrl_oracle as O  # noqa: E402

CONFIGS = {"per_leg": {}, "norm": {"norm_reward": True}, "global": {"global_reward": True}}
BASE_CFG = {"ctrl_cost_weight": 0.25, "contact_cost_weight": 0.025}
STEPS = 5


def lift_with_constructors(mod):
    ns = MA.lift_module(mod)
    tree = MA.ast.parse(open(f"{MA.REF}/simulation_envs/quantruped_adaptor_multi_environment.py").read())
    root = next(n for n in tree.body if isinstance(n, MA.ast.ClassDef) and n.name == "QuantrupedMultiPoliciesEnv")
    init = next(n for n in root.body if isinstance(n, MA.ast.FunctionDef) and n.name == "__init__")
    env_ns = {"np": np, "random": random, "Iterable": Iterable}
    exec(compile(MA.ast.fix_missing_locations(MA.ast.Module(body=[init], type_ignores=[])), "adaptor.__init__", "exec"), env_ns)
    ns["QuantrupedMultiPoliciesEnv"].__init__ = env_ns["__init__"]          # the reference's own root constructor again

    filters = {}

    class MeanStdFilterSingleton:          # simulation_envs/observation_filter.py:3-12
        @staticmethod
        def get_instance(shape):
            return filters.setdefault(tuple(shape), O.MeanStdFilter(tuple(shape)))
    ns["MeanStdFilterSingleton"] = MeanStdFilterSingleton
    ns["MeanStdFilter"] = O.MeanStdFilter
    return ns, filters


def make_sim(script):
    sim = MA.sim_helpers()
    sim.t, sim.seen = 0, []
    sim.sim = types.SimpleNamespace(data=types.SimpleNamespace(cfrc_ext=None))
    sim.action_space = sim.observation_space = None
    sim.scale_mass = lambda f: None
    sim.set_target_velocity = lambda v: None

    def step(action):
        sim.seen.append(np.array(action, dtype=np.float64))
        t = sim.t
        sim.sim.data.cfrc_ext = script["cfrc"][t]
        sim.t += 1
        return script["obs_full"][t].copy(), 0.0, False, {"reward_forward": script["fw"][t]}
    sim.step = step
    return sim


def main():
    rng = np.random.default_rng(77)
    script = {"obs_full": rng.standard_normal((STEPS, 43)) * rng.uniform(0.2, 3.0, 43),
              "fw": rng.standard_normal(STEPS), "cfrc": 2.0 * rng.standard_normal((STEPS, 14, 6))}
    script["obs_full"][:, 1:5] /= np.linalg.norm(script["obs_full"][:, 1:5], axis=1, keepdims=True)
    script["cfrc"][rng.random((STEPS, 14, 6)) < 0.5] = 0.0
    out = dict(obs_full=script["obs_full"], fw=script["fw"], cfrc=script["cfrc"])
    for scope, (mod, cname) in sorted(MA.scope_table().items()):
        ns, filters = lift_with_constructors(mod)
        cls = ns[cname]
        agents = list(cls.agent_names)
        A = 8 // len(agents)
        actions = np.clip(1.2 * rng.standard_normal((STEPS, len(agents), A)), -1.0, 1.0)
        out[f"actions/{scope}"] = actions
        for tag, cfg in CONFIGS.items():
            filters.clear()
            sim = make_sim(script)
            ns["QuantrupedMultiPoliciesEnv"].create_env = lambda self, use_target_velocity=False, sim=sim, **kw: (
                [setattr(sim, k, v) for k, v in kw.items()], sim)[1]
            env = cls(dict(BASE_CFG, **cfg))
            assert sim.ctrl_cost_weight == 0.25 and sim.contact_cost_weight == 0.025
            out[f"{scope}/{tag}/reward_fn"] = np.array(env.distribute_reward.__func__.__name__)
            out[f"{scope}/{tag}/normalize_rewards"] = np.array(bool(env.normalize_rewards))
            rews, obss = [], []
            for t in range(STEPS):
                obs_d, rew_d, done, info = env.step({a: actions[t, i] for i, a in enumerate(agents)})
                assert done == {"__all__": False} and info == {}
                rews.append([rew_d[a] for a in agents])
                obss.append([np.asarray(obs_d[a][1] if isinstance(obs_d[a], tuple) else obs_d[a], dtype=np.float64) for a in agents])
                for i, a in enumerate(agents):
                    if isinstance(obs_d[a], tuple):
                        assert int(obs_d[a][0][0]) == i
            out[f"{scope}/{tag}/sim_action"] = np.stack(sim.seen)
            out[f"{scope}/{tag}/rew"] = np.asarray(rews, dtype=np.float64)
            out[f"{scope}/{tag}/obs"] = np.asarray(obss, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "env_step.npz"), **out)
    fns = sorted({(k.split("/")[0], k.split("/")[1], str(v)) for k, v in out.items() if k.endswith("reward_fn")})
    print("wrote env_step.npz;", len(out), "arrays; reward functions in use:", sorted({f for _, _, f in fns}))
    print([x for x in fns if "GlobalCost" in x[0]])


if __name__ == "__main__":
    main()
