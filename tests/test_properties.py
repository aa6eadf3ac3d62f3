"""Property tests (hypothesis) of the oracle's streaming pieces — size-independent invariants the GPU parity tests
rely on: Chan-merge associativity of the filter statistics, GAE columns == per-fragment scipy procedure for arbitrary
done patterns, standardisation idempotence, KL-coefficient rule, checkpoint round trip for arbitrary shapes."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle.ddrl_oracle as O

_S = dict(max_examples=40, deadline=None)


@settings(**_S)
@given(st.integers(2, 120), st.integers(1, 6), st.data())
def test_filter_merge_is_split_independent(n, d, data):
    """Any split of the rows into blocks, merged in any grouping, gives the sequential Welford state (the multi-GPU
    merge of per-rank partials relies on this)."""
    seed = data.draw(st.integers(0, 2 ** 16))
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)) * rng.uniform(0.1, 50.0, size=d) + rng.uniform(-5, 5, size=d)
    seq = O.RunningStat((d,))
    for r in x:
        seq.push(r)
    cuts = sorted(set(data.draw(st.lists(st.integers(1, n - 1), min_size=1, max_size=5))))
    blocks = [b for b in np.split(x, cuts) if len(b)]
    left = O.batch_stat(blocks[0])
    for b in blocks[1:]:
        left.update(O.batch_stat(b))
    right = O.batch_stat(blocks[-1])
    for b in reversed(blocks[:-1]):
        tmp = O.batch_stat(b)
        tmp.update(right)
        right = tmp
    for m in (left, right):
        assert m.n == seq.n == n
        np.testing.assert_allclose(m.mean, seq.mean, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(m._S, seq._S, rtol=1e-9, atol=1e-9)


@settings(**_S)
@given(st.integers(1, 40), st.integers(1, 5), st.floats(0.0, 0.5), st.integers(0, 2 ** 16))
def test_gae_columns_equal_fragment_procedure_for_any_done_pattern(T, C, pdone, seed):
    rng = np.random.default_rng(seed)
    r = rng.standard_normal((T, C)).astype(np.float32)
    v = (5 * rng.standard_normal((T, C))).astype(np.float32)
    d = (rng.random((T, C)) < pdone).astype(np.uint8)
    vb = rng.standard_normal(C).astype(np.float32)
    a1, t1 = O.gae_columns(r, v, d, vb)            # per fragment: scipy.signal.lfilter like RLlib
    a2, t2 = O.gae_recurrence(r, v, d, vb)         # masked recurrence the CUDA kernel implements
    np.testing.assert_allclose(a1, a2, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(t1, t2, rtol=2e-6, atol=2e-6)
    # a done at the last step makes the bootstrap value irrelevant
    d2 = d.copy(); d2[-1] = 1
    a3, _ = O.gae_columns(r, v, d2, vb)
    a4, _ = O.gae_columns(r, v, d2, vb + 100.0)
    np.testing.assert_array_equal(a3, a4)


@settings(**_S)
@given(st.integers(8, 500), st.floats(0.1, 100.0), st.floats(-10, 10), st.integers(0, 2 ** 16))
def test_standardisation_is_idempotent_and_scale_free(n, scale, shift, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(n).astype(np.float32)
    s1 = O.standardized(a)
    assert abs(float(s1.mean())) < 1e-4 and abs(float(s1.std()) - 1.0) < 1e-3
    np.testing.assert_allclose(O.standardized(s1), s1, rtol=1e-4, atol=1e-5)
    # affine invariance (within float32 cancellation of the shift; the 1e-4 floor on the std is far away at these scales)
    np.testing.assert_allclose(O.standardized((a * np.float32(scale) + np.float32(shift)).astype(np.float32)), s1, rtol=0, atol=2e-2)


@settings(**_S)
@given(st.floats(1e-4, 10.0), st.floats(0.0, 1.0))
def test_kl_coefficient_rule(c, kl):
    out = O.update_kl(c, kl, 0.01)
    assert out == (c * 1.5 if kl > 0.02 else c * 0.5 if kl < 0.005 else c)


@settings(max_examples=15, deadline=None)
@given(st.integers(1, 46), st.sampled_from([1, 2, 4, 8]), st.integers(1, 4), st.integers(0, 2 ** 16))
def test_checkpoint_round_trip_any_shape(D, A, P, seed):
    import tempfile
    from collections import OrderedDict
    from ddrl_b200 import checkpoint as C
    rng = np.random.default_rng(seed)
    shapes = [(D, 64), (64,), (D, 64), (64,), (64, 64), (64,), (64, 64), (64,), (64, 2 * A), (2 * A,), (64, 1), (1,)]
    NP = 128 * D + 130 * A + 8513
    pols = OrderedDict()
    for p in range(P):
        pols[f"policy_{p}"] = C.PolicyCheckpoint(
            theta=rng.standard_normal(NP).astype(np.float32), shapes=shapes, adam_m=rng.standard_normal(NP).astype(np.float32),
            adam_v=rng.random(NP).astype(np.float32), beta_powers=rng.random(2).astype(np.float32), filter_n=int(rng.integers(2, 10 ** 9)),
            filter_M=rng.standard_normal(D), filter_S=rng.random(D) * 1e6, learner_stats={"cur_kl_coeff": 0.3, "kl": 0.01})
    ck = C.Checkpoint(policies=pols, counters={"num_steps_trained": int(rng.integers(0, 10 ** 8))})
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ck")
        C.save_rllib_checkpoint(path, ck)
        back = C.load_rllib_checkpoint(path)
    assert list(back.policies) == list(pols) and back.counters == ck.counters
    for pid, pc in pols.items():
        b = back.policies[pid]
        assert np.array_equal(b.theta, pc.theta) and np.array_equal(b.adam_m, pc.adam_m) and np.array_equal(b.adam_v, pc.adam_v)
        assert b.filter_n == pc.filter_n and np.array_equal(b.filter_M, pc.filter_M) and np.array_equal(b.filter_S, pc.filter_S)
        assert b.obs_dim == D and b.act_dim == A
