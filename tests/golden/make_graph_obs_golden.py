#!/usr/bin/env python
"""Golden vectors for the shared-graph env's node features (leg quaternion encoding + obs routing).

Run HERE (container with /root/reference mounted), never on the GPU box:

    python tests/golden/make_graph_obs_golden.py

The reference module cannot be imported (gym, ray and MuJoCo are absent), so the four pure-numpy methods of
`QuantrupedDecentralizedSharedGraphEnv` (simulation_envs/quantruped_GraphDecentralizedController_environments.py:
`leg_encoding`, `quaternion_multiply`, `leg_encoding_ego`, `distribute_observations`) and its `leg_angles` table are
lifted from the reference file's AST and executed UNMODIFIED on a stub `self` that supplies what `__init__` would
have built: `obs_indices`, `agent_names`, `adj`, and
`_normalize_observation` = the oracle's MeanStdFilter restatement (RLlib's filter is not installed here).
`obs_indices` come from the reference's own `QuAntrupedEnv.get_obs_indices` + `OBS_FIELDS` (simulation_envs/
quantruped_v3.py:68-95,282-300), lifted the same way (the index lists in the comments at :27-28 of the graph env are
stale: they are sorted, the code is prefix-major).

Writes tests/golden/graph_obs.npz:
    obs_full [64, 43] f64      seeded synthetic observations (unit body quaternion in columns 1:5)
    seq      [64, 4, 23] f64   outputs with the filter UPDATING on every call, as the env does (sequential)
    frozen   [64, 4, 23] f64   outputs with the filter frozen at its state after the 64 pushes (what a batched call sees)
    mean, std [43] f64         that frozen state
    table    [4, 19] i32       the reference's obs_indices (agent order FL, HL, HR, FR)
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ddrl_oracle as O  # noqa: E402

SRC = "/root/reference/simulation_envs/quantruped_GraphDecentralizedController_environments.py"
SIM = "/root/reference/simulation_envs/quantruped_v3.py"
WANT = {"leg_encoding", "quaternion_multiply", "leg_encoding_ego", "distribute_observations"}


def lift(src, cls_name, funcs, assigns):
    """Class `cls_name` of file `src` reduced to the named methods and class-level assignments, compiled unmodified."""
    tree = ast.parse(open(src).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    body = [n for n in cls.body if (isinstance(n, ast.FunctionDef) and n.name in funcs)
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in assigns)]
    assert {n.name for n in body if isinstance(n, ast.FunctionDef)} == set(funcs)
    mod = ast.Module(body=[ast.ClassDef(name="Lifted", bases=[], keywords=[], body=body, decorator_list=[])], type_ignores=[])
    ns = {"np": np}
    exec(compile(ast.fix_missing_locations(mod), src, "exec"), ns)
    return ns["Lifted"]


def main():
    env = lift(SRC, "QuantrupedDecentralizedSharedGraphEnv", WANT, {"leg_angles"})()
    sim = lift(SIM, "QuAntrupedEnv", {"get_obs_indices"}, {"OBS_FIELDS"})()
    env.agent_names = ["agent_FL", "agent_HL", "agent_HR", "agent_FR"]
    env.obs_indices = {a: [int(i) for i in sim.get_obs_indices(["body", a[-2:].lower()])] for a in env.agent_names}   # ref :26-31
    for a in env.agent_names:      # the oracle's table restatement equals the reference's code
        assert env.obs_indices[a] == O.get_obs_indices(["body", a[-2:].lower()]) and len(env.obs_indices[a]) == 19
    env.adj = O.ring_adjacency().numpy().astype(np.float64)

    rng = np.random.default_rng(20260101)
    obs = rng.standard_normal((64, 43)) * rng.uniform(0.1, 3.0, 43) + rng.uniform(-1, 1, 43)
    q = rng.standard_normal((64, 4))
    obs[:, 1:5] = q / np.linalg.norm(q, axis=1, keepdims=True)

    filt = O.MeanStdFilter((43,))
    env._normalize_observation = lambda o: filt(o)                       # updates on every call, like the env
    seq = []
    for t in range(64):
        d = env.distribute_observations(obs[t])
        assert list(d) == env.agent_names and all(int(d[a][0][0]) == i for i, a in enumerate(env.agent_names))
        assert all(np.array_equal(d[a][1], d["agent_FL"][1]) and d[a][2] is env.adj for a in d)
        seq.append(d["agent_FL"][1])
    env._normalize_observation = lambda o: filt(o, update=False)         # frozen statistics
    frozen = [env.distribute_observations(obs[t])["agent_FL"][1] for t in range(64)]
    np.savez_compressed(os.path.join(HERE, "graph_obs.npz"), obs_full=obs, seq=np.stack(seq), frozen=np.stack(frozen),
                        mean=filt.rs.mean, std=filt.rs.std,
                        table=np.asarray([env.obs_indices[a] for a in env.agent_names], dtype=np.int32))
    print("wrote graph_obs.npz", np.stack(seq).shape)


if __name__ == "__main__":
    main()
