#!/usr/bin/env python
"""Golden vectors for the terrain curriculum callback.

    python tests/golden/make_curriculum_golden.py        (HERE only: needs /root/reference)

`QuantrupedMultiPoliciesEnv.update_environment_after_epoch` / `update_after_epoch`
(simulation_envs/quantruped_adaptor_multi_environment.py:94-122) are lifted from the reference file's AST and run
unmodified on a stub whose simulator records the smoothness it is given; `np.random.seed(7)` fixes the draws.
The schedule is experiment 3's (`range_smoothness=[1., 0.6]`, `range_last_timestep` stretched to 10 M so that both branches are hit
(train_experiment_3_architecture_curriculum_targetvel.py:101-102 uses 4 M).  Writes tests/golden/curriculum.json."""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_graph_obs_golden import lift  # noqa: E402

SRC = "/root/reference/simulation_envs/quantruped_adaptor_multi_environment.py"


def main():
    Env = lift(SRC, "QuantrupedMultiPoliciesEnv", {"update_environment_after_epoch", "update_after_epoch"}, set())
    seen = []
    env = Env()
    env.curriculum_learning = True
    env.curriculum_initial_smoothness, env.curriculum_target_smoothness = 1.0, 0.6
    env.current_smoothness = 1.0
    env.curriculum_last_timestep = 10_000_000
    env.env = types.SimpleNamespace(set_hf_parameter=seen.append, create_new_random_hfield=lambda: None, reset=lambda: None)
    np.random.seed(7)
    ts = [16_000 * k for k in (1, 10, 100, 300, 600, 624, 625, 626, 900, 1250)]
    for t in ts:
        env.update_environment_after_epoch(t)
    assert len(seen) == len(ts)
    json.dump({"seed": 7, "range_smoothness": [1.0, 0.6], "range_last_timestep": 10_000_000, "timesteps_total": ts,
               "smoothness": [float(s) for s in seen]}, open(os.path.join(HERE, "curriculum.json"), "w"), indent=1)
    print("wrote curriculum.json", seen[:3])


if __name__ == "__main__":
    main()
