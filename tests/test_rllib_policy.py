"""CPU: host logic of the RLlib-facing learner surface — batch stacking and validation, the random
draws of the minibatch order, the ragged-tail rule, the curriculum schedule against the reference's own code."""
import json
import os

import numpy as np
import pytest

from ddrl_b200 import rllib_policy as RP
from tests.util import GOLDEN

NAMES = ["policy_FL", "policy_HL", "policy_HR", "policy_FR"]
D, A = 19, 2


def _batches(R=12, seed=0):
    rng = np.random.default_rng(seed)
    return {p: {RP.OBS: rng.standard_normal((R, D)), RP.ACTIONS: rng.standard_normal((R, A)),
                RP.ACTION_DIST_INPUTS: rng.standard_normal((R, 2 * A)), RP.ACTION_LOGP: rng.standard_normal(R),
                RP.VF_PREDS: rng.standard_normal(R), RP.ADVANTAGES: rng.standard_normal(R),
                RP.VALUE_TARGETS: rng.standard_normal(R)} for p in reversed(NAMES)}       # dict order != policy order


def test_stack_orders_by_policy_names_and_casts_to_float32():
    b = _batches()
    s = RP.stack_policy_batches(b, NAMES, D, A)
    assert set(s) == set(RP.TRAIN_COLUMNS)
    for col, arr in s.items():
        assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"] and arr.shape[:2] == (4, 12)
        for i, p in enumerate(NAMES):
            assert np.array_equal(arr[i], np.asarray(b[p][col], dtype=np.float32))


@pytest.mark.parametrize("breaker,msg", [
    (lambda b: b.pop("policy_HL"), "lacks"),
    (lambda b: b.update(policy_XX=b["policy_FL"]), "unknown policy"),
    (lambda b: b["policy_HR"].pop(RP.ADVANTAGES), "lacks column"),
    (lambda b: b["policy_HR"].update({RP.OBS: b["policy_HR"][RP.OBS][:, :18]}), "expected shape"),
    (lambda b: b["policy_FR"].update({RP.VF_PREDS: b["policy_FR"][RP.VF_PREDS][:7]}), "equally long"),
    (lambda b: b["policy_FL"][RP.ACTION_LOGP].__setitem__(3, np.nan), "non-finite"),
    (lambda b: [v.update({k: a[:0] for k, a in v.items()}) for v in b.values()], "empty"),
])
def test_stack_rejects_malformed_batches(breaker, msg):
    b = _batches()
    breaker(b)
    with pytest.raises(RP.BatchError, match=msg):
        RP.stack_policy_batches(b, NAMES, D, A)


def test_minibatch_order_replays_numpy_random_state():
    P, R, E, nb = 3, 40, 4, 5
    shuffle, perms = RP.draw_minibatch_order(np.random.RandomState(5), P, R, E, nb)
    rng = np.random.RandomState(5)
    want_shuffle = [rng.permutation(R) for _ in range(P)]
    want_perms = [[rng.permutation(nb) for _ in range(E)] for _ in range(P)]
    assert shuffle.dtype == perms.dtype == np.int32 and shuffle.shape == (P, R) and perms.shape == (P, E, nb)
    assert np.array_equal(shuffle, want_shuffle) and np.array_equal(perms, want_perms)
    none, perms2 = RP.draw_minibatch_order(np.random.RandomState(5), P, R, E, nb, shuffle_sequences=False)
    assert none is None and all(sorted(p) == list(range(nb)) for pp in perms2 for p in pp)


@pytest.mark.parametrize("R,mb,want", [(16000, 128, (16000, 125)), (520, 128, (512, 4)), (100, 128, (100, 1)), (128, 128, (128, 1))])
def test_usable_rows_keeps_whole_minibatches(R, mb, want):
    assert RP.usable_rows(R, mb) == want


def test_curriculum_schedule_equals_the_reference_callback():
    g = json.load(open(os.path.join(GOLDEN, "curriculum.json")))
    rng = np.random.RandomState(g["seed"])
    lo, hi = g["range_smoothness"]
    got = [RP.curriculum_smoothness(t, lo, hi, g["range_last_timestep"], rng.rand()) for t in g["timesteps_total"]]
    assert got == g["smoothness"]                      # float64 on both sides, same operation order: bit equal
    assert all(hi <= s <= lo for s in got)


def test_rllib_surface_orchestration_with_oracle_mocked_kernels():
    """tests/host_dryrun.py: the GPU tests of the RLlib surface executed on CPU with every kernel replaced by the oracle —
    host orchestration only (own process: it monkeypatches torch)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "host_dryrun.py")], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert out.stdout.split() == ["test1", "fp32", "ok", "test2", "fp32", "ok", "test1", "tc", "ok", "test2", "tc", "ok"]
