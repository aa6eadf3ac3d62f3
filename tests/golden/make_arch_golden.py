#!/usr/bin/env python
"""Golden table of the multi-agent architectures from the reference's own classes.

    python tests/golden/make_arch_golden.py        (HERE only: needs /root/reference)

The env modules import gym / ray / MuJoCo, so only their STATIC interface is lifted from the ASTs and executed unmodified:
class-level constants (`policy_names`, `agent_names`, `leg_angles`) and the `@staticmethod`s `policy_mapping_fn` /
`return_policies`, with the class hierarchy kept (the root `MultiAgentEnv` becomes `object`) and `gym.spaces` replaced by
`ddrl_b200.spaces` (same constructor signatures).  The `--policy_scope` -> class table is read from the if-chain of
train_experiment_1_architecture_on_flat.py:63-90.  Writes tests/golden/architectures.json:
    {scope: {"class": name, "policy_names": [...], "agent_names": [...], "mapping": {agent_id: policy_id},
             "policies": {"flat" | "tvel": {policy_id: {"obs": <space>, "act": <space>}}},
             "tables" | "tables_tvel": {"obs": {agent: [index...]}, "act": {agent: [index...]},
                                        "contact": {agent: [[row...], [weight...]]}}}}
with <space> = ["Box", shape, dtype] | ["MultiDiscrete", nvec] | ["Tuple", [<space>, ...]].
"tables" / "tables_tvel" are what each class's CONSTRUCTOR builds (43 fields / the 44 fields of the target-velocity env): the constructors are executed too, on
a root class whose `__init__` only provides `self.env` = the index helpers of `QuAntrupedEnv` (quantruped_v3.py:68-112,282-341,
lifted the same way) instead of creating the MuJoCo simulation."""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ddrl_b200 import spaces  # noqa: E402

REF = "/root/reference"
PROBE_IDS = ["agent_FL", "agent_HL", "agent_HR", "agent_FR", "agent_LEFT", "agent_RIGHT", "agent_FLHR", "agent_HLFR",
             "central_agent", "agent_FL_1", "agent_HR7", "somebody_else"]


def static_part(cls: ast.ClassDef) -> ast.ClassDef:
    """Class-level constants and all methods (only the static ones and the constructors are ever called)."""
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef)
            or (isinstance(n, ast.Assign) and all(isinstance(t, ast.Name) for t in n.targets))]
    bases = [b if getattr(b, "id", "") != "MultiAgentEnv" else ast.Name(id="object", ctx=ast.Load()) for b in cls.bases]
    return ast.ClassDef(name=cls.name, bases=bases, keywords=[], body=keep or [ast.Pass()], decorator_list=[])


def sim_helpers(tvel: bool = False):
    """`QuAntrupedEnv` reduced to its field lists and index helpers; `tvel`: with the 44-entry OBS_FIELDS of
    `QuAntrupedTVelEnv` as the class declares it (quantruped_v3.py:354-383).  (At run time `set_target_velocity` appends the
    target-velocity field a SECOND time (:394-400), which contradicts `return_policies` (D + 1) and the published TVel
    checkpoints (D + 1); the declared list is the one both agree with.)"""
    tree = ast.parse(open(f"{REF}/simulation_envs/quantruped_v3.py").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "QuAntrupedEnv")
    want = {"get_obs_indices", "get_action_indices", "get_contact_force_indices"}
    body = [n for n in cls.body if (isinstance(n, ast.FunctionDef) and n.name in want)
            or (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "").endswith("_FIELDS"))]
    if tvel:
        tv = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "QuAntrupedTVelEnv")
        body += [n for n in tv.body if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "OBS_FIELDS"]   # overrides
    ns = {"np": np}
    mod = ast.Module(body=[ast.ClassDef(name="Sim", bases=[], keywords=[], body=body, decorator_list=[])], type_ignores=[])
    exec(compile(ast.fix_missing_locations(mod), "quantruped_v3.py", "exec"), ns)
    return ns["Sim"]()


def lift_module(name: str):
    """Static interface + constructors of simulation_envs/<name>.py in its own namespace (two modules define a class of the
    same name)."""
    class MeanStdFilter:                      # the graph env builds one in its constructor; never called here
        def __init__(self, shape):
            self.shape = shape
    ns = {"np": np, "spaces": spaces, "MeanStdFilter": MeanStdFilter}
    files = [f"{REF}/simulation_envs/quantruped_adaptor_multi_environment.py"]
    if name != "quantruped_adaptor_multi_environment":
        files.append(f"{REF}/simulation_envs/{name}.py")
    for f in files:
        tree = ast.parse(open(f).read())
        body = [static_part(n) for n in tree.body if isinstance(n, ast.ClassDef)]
        exec(compile(ast.fix_missing_locations(ast.Module(body=body, type_ignores=[])), f, "exec"), ns)

    def root_init(self, config):              # instead of QuantrupedMultiPoliciesEnv.__init__ (creates the simulation)
        self.env = sim_helpers(tvel=bool(config.get("target_velocity")))
    ns["QuantrupedMultiPoliciesEnv"].__init__ = root_init
    return ns


def tables(env):
    def cfi(v):
        idx, w = v
        return [np.asarray(idx).astype(int).tolist(), np.asarray(w, dtype=float).reshape(-1).tolist()]
    out = {}
    if hasattr(env, "obs_indices"):
        out["obs"] = {a: np.asarray(v).astype(int).tolist() for a, v in env.obs_indices.items()}
    if hasattr(env, "action_indices"):
        out["act"] = {a: np.asarray(v).astype(int).tolist() for a, v in env.action_indices.items()}
    if hasattr(env, "contact_force_indices"):
        out["contact"] = {a: cfi(v) for a, v in env.contact_force_indices.items()}
    return out


def scope_table():
    """{policy_scope: (module, class name)} from the import if-chain of the training script."""
    tree = ast.parse(open(f"{REF}/train_experiment_1_architecture_on_flat.py").read())
    out = {}

    def walk(node):
        if isinstance(node, ast.If) and isinstance(node.test, ast.Compare) and getattr(node.test.left, "id", "") == "policy_scope":
            scope = node.test.comparators[0].value
            imp = next(n for n in node.body if isinstance(n, ast.ImportFrom))
            out[scope] = (imp.module.split(".")[-1], imp.names[0].name)
            for n in node.orelse:
                if isinstance(n, ast.If):
                    walk(n)
                elif isinstance(n, ast.ImportFrom):
                    out["QuantrupedMultiEnv_Centralized"] = (n.module.split(".")[-1], n.names[0].name)   # else branch (:61, :89-90)
    for n in tree.body:
        walk(n)
    return out


def describe(sp):
    if isinstance(sp, spaces.Tuple):
        return ["Tuple", [describe(s) for s in sp]]
    if isinstance(sp, spaces.MultiDiscrete):
        return ["MultiDiscrete", np.asarray(sp.nvec).astype(int).tolist()]
    return ["Box", list(sp.shape), str(np.dtype(sp.dtype))]


def main():
    table, out = scope_table(), {}
    for scope, (mod, cname) in sorted(table.items()):
        cls = lift_module(mod)[cname]
        pol = {}
        for tag, tv in (("flat", False), ("tvel", True)):
            pol[tag] = {pid: {"obs": describe(spec[1]), "act": describe(spec[2])}
                        for pid, spec in cls.return_policies(use_target_velocity=tv).items()}
        out[scope] = {"class": cname, "policy_names": list(cls.policy_names), "agent_names": list(getattr(cls, "agent_names", [])),
                      "mapping": {a: cls.policy_mapping_fn(a) for a in PROBE_IDS}, "policies": pol, "tables": tables(cls({})),
                      "tables_tvel": tables(cls({"target_velocity": [1.0]}))}
    json.dump(out, open(os.path.join(HERE, "architectures.json"), "w"), indent=1, sort_keys=True)
    print("wrote architectures.json:", len(out), "scopes")


if __name__ == "__main__":
    main()
