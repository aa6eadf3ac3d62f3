"""Shared helpers for the tests (golden fixture loading, synthetic rollouts).  Test infrastructure only."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# (arch file stem) -> (policy order used by the env's policy_names, D, A)
ARCHS = {
    "Centralized": (["central_policy"], 43, 8),
    "FullyDecentral": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 19, 2),
    "Local": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 35, 2),
    "SingleDiagonal": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 27, 2),
    "SingleNeighbor": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 27, 2),
    "SingleToFront": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 27, 2),
    "TwoDiags": (["policy_FLHR", "policy_HLFR"], 27, 4),
    "TwoSides": (["policy_LEFT", "policy_RIGHT"], 27, 4),
    "Centralized_TVel": (["central_policy"], 44, 8),
    "FullyDecentral_TVel": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 20, 2),
    "Local_TVel": (["policy_FL", "policy_HL", "policy_HR", "policy_FR"], 36, 2),
    "TwoSides_TVel": (["policy_LEFT", "policy_RIGHT"], 28, 4),
}


def load_ckpt(arch):
    return np.load(os.path.join(GOLDEN, f"ckpt_{arch}.npz"))


def ckpt_theta(arch):
    """-> theta [P, NP] float32 in the env's policy order, filters list of (n, M, S)."""
    pids, D, A = ARCHS[arch]
    z = load_ckpt(arch)
    theta = np.stack([z[f"{pid}/theta"] for pid in pids])
    filt = [(int(z[f"{pid}/filter_n"]), z[f"{pid}/filter_M"], z[f"{pid}/filter_S"]) for pid in pids]
    return theta, filt, D, A


def learner_stats():
    return json.load(open(os.path.join(GOLDEN, "learner_stats.json")))


def synth_obs(filt, R, seed):
    """raw observations ~ N(mean, std^2) of a checkpoint filter -> [P, R, D] float32."""
    rng = np.random.default_rng(seed)
    out = []
    for n, M, S in filt:
        std = np.sqrt(S / (n - 1))
        out.append((M + std * rng.standard_normal((R, M.shape[0]))).astype(np.float32))
    return np.stack(out)


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| : error relative to the tensor's scale (robust for entries near zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30)) if a.size else 0.0


def t64(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float64))
