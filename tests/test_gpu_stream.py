"""GPU parity of the streaming kernels: MeanStdFilter, GAE, standardisation, shuffle, DiagGaussian sample,
LegCoupling — CUDA (C ABI) vs the oracle restating RLlib 1.0.1."""
import numpy as np
import pytest
import torch

from tests.util import ckpt_theta, scaled_err, synth_obs

pytestmark = pytest.mark.gpu


def _O():
    import oracle.ddrl_oracle as O
    return O


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


@pytest.mark.parametrize("R,dtype", [(1, np.float32), (2, np.float32), (127, np.float32), (128, np.float64),
                                     (1000, np.float32), (20000, np.float32), (3001, np.float64)])
def test_filter_update_matches_sequential_welford(R, dtype):
    """count bit-exact; mean / M2 within 1e-12 rel of the reference's sequential push order; two successive
    updates (state carried over) and the refreshed normalisation constants."""
    from ddrl_b200 import kernels as K
    O = _O()
    _, filt, D, _ = ckpt_theta("FullyDecentral")
    P = len(filt)
    x1 = synth_obs(filt, R, 1).astype(dtype)
    x2 = synth_obs(filt, R // 2 + 1, 2).astype(dtype)
    n = torch.zeros(P, dtype=torch.int64, device="cuda")
    M = torch.zeros(P, D, dtype=torch.float64, device="cuda")
    S = torch.zeros(P, D, dtype=torch.float64, device="cuda")
    refs = [O.MeanStdFilter((D,), clip=None) for _ in range(P)]
    for x in (x1, x2):
        norm = K.filter_update(_dev(x), n, M, S)
        torch.cuda.synchronize()
        for p in range(P):
            refs[p](x[p])
            rs = refs[p].rs
            assert int(n[p]) == rs._n                                   # bit-exact count
            np.testing.assert_allclose(M[p].cpu().numpy(), rs._M, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(S[p].cpu().numpy(), rs._S, rtol=1e-11, atol=1e-18)
            np.testing.assert_allclose(norm[p, 0].cpu().numpy(), rs.mean, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(norm[p, 1].cpu().numpy(), 1.0 / (rs.std + 1e-8), rtol=1e-11)


def test_filter_partial_merge_equals_single_update():
    """Data-parallel path: per-rank partials concatenated in rank order == one update over all rows."""
    from ddrl_b200 import kernels as K
    _, filt, D, _ = ckpt_theta("Local")
    P = len(filt)
    x = synth_obs(filt, 3000, 3)
    st = lambda: (torch.zeros(P, dtype=torch.int64, device="cuda"), torch.zeros(P, D, dtype=torch.float64, device="cuda"),
                  torch.zeros(P, D, dtype=torch.float64, device="cuda"))
    n1, M1, S1 = st()
    K.filter_update(_dev(x), n1, M1, S1)
    n2, M2, S2 = st()
    parts = [K.filter_partial(_dev(x[:, i * 750:(i + 1) * 750])) for i in range(4)]
    K.filter_merge(torch.cat(parts, dim=1).contiguous(), 3000, n2, M2, S2)
    torch.cuda.synchronize()
    assert torch.equal(n1, n2)
    np.testing.assert_allclose(M2.cpu().numpy(), M1.cpu().numpy(), rtol=1e-13)
    np.testing.assert_allclose(S2.cpu().numpy(), S1.cpu().numpy(), rtol=1e-12)


@pytest.mark.parametrize("T,C,cpe,pdone", [(1, 1, 1, 0.0), (32, 300, 1, 0.03), (200, 8, 1, 0.005), (32, 64, 4, 0.1),
                                           (7, 1000, 2, 1.0), (50, 33, 1, 0.0)])
def test_gae_matches_reference_fragment_procedure(T, C, cpe, pdone):
    """vs compute_advantages applied per fragment (scipy.signal.lfilter in float64, cast to float32)."""
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(T * 1000 + C)
    P = 2
    r = (0.3 + 0.5 * rng.standard_normal((P, T, C))).astype(np.float32)
    v = (20 * rng.standard_normal((P, T, C))).astype(np.float32)
    vb = (20 * rng.standard_normal((P, C))).astype(np.float32)
    d_env = (rng.random((T, C // cpe)) < pdone).astype(np.uint8)
    adv, vt, mom = K.gae(_dev(r), _dev(v), _dev(d_env), _dev(vb), cpe, 0.99, 0.95)
    torch.cuda.synchronize()
    d_col = np.repeat(d_env, cpe, axis=1)
    for p in range(P):
        a_ref, vt_ref = O.gae_columns(r[p], v[p], d_col, vb[p], 0.99, 0.95)
        # float64 recursion on both sides, one rounding to float32: allow 1 ulp
        np.testing.assert_allclose(adv[p].cpu().numpy(), a_ref, rtol=2e-7, atol=1e-6)
        np.testing.assert_allclose(vt[p].cpu().numpy(), vt_ref, rtol=2e-7, atol=1e-6)
        a64 = adv[p].cpu().numpy().astype(np.float64)
        assert mom[p, 0].item() == T * C
        np.testing.assert_allclose(mom[p, 1].item(), a64.sum(), rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(mom[p, 2].item(), (a64 ** 2).sum(), rtol=1e-12)


def test_adv_standardize_matches_numpy():
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(5)
    a = (3.0 + 7.0 * rng.standard_normal((3, 4096))).astype(np.float32)
    a[2] = 1.25                                # zero variance -> max(1e-4, std) floor
    mom = np.stack([[a.shape[1], a[p].astype(np.float64).sum(), (a[p].astype(np.float64) ** 2).sum()] for p in range(3)])
    out = K.adv_standardize(_dev(a.copy()), _dev(mom))
    torch.cuda.synchronize()
    for p in range(3):
        ref = O.standardized(a[p])
        np.testing.assert_allclose(out[p].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)


def test_gather_rows_is_exact():
    from ddrl_b200 import kernels as K
    rng = np.random.default_rng(6)
    for W in (1, 2, 19, 44):
        src = rng.standard_normal((3, 1000, W)).astype(np.float32)
        perm = np.stack([rng.permutation(1000) for _ in range(3)]).astype(np.int32)
        dst = K.gather_rows(_dev(src), _dev(perm))
        torch.cuda.synchronize()
        ref = np.stack([src[p][perm[p]] for p in range(3)])
        assert np.array_equal(dst.cpu().numpy(), ref)


def test_dg_sample_and_leg_coupling():
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(8)
    for A in (2, 4, 8):
        lg = rng.standard_normal((500, 2 * A)).astype(np.float32)
        eps = rng.standard_normal((500, A)).astype(np.float32)
        act, logp = K.dg_sample(_dev(lg), _dev(eps))
        a_ref = O.dg_sample(torch.from_numpy(lg).double(), torch.from_numpy(eps).double())
        assert scaled_err(act.cpu().numpy(), a_ref.numpy()) < 1e-6
        assert scaled_err(logp.cpu().numpy(), O.dg_logp(torch.from_numpy(lg).double(), a_ref).numpy()) < 1e-5
    lg = rng.standard_normal((333, 4)).astype(np.float32)
    nid = rng.integers(0, 4, size=333).astype(np.int32)
    coup = np.array(O.COUPLING_INIT, dtype=np.float32) * 0.7
    out = K.leg_coupling_(_dev(lg.copy()), _dev(nid), _dev(coup))
    ref = O.leg_coupling(torch.from_numpy(lg), torch.from_numpy(nid), torch.from_numpy(coup))
    assert np.array_equal(out.cpu().numpy(), ref.numpy())


def test_bad_arguments_fail_loudly():
    from ddrl_b200 import kernels as K
    from ddrl_b200._lib import DDRLError
    with pytest.raises(DDRLError):
        K.fcnet_forward(torch.zeros(1, 10, device="cuda"), torch.zeros(1, 4, 19, device="cuda"), 2)   # wrong NP
    with pytest.raises(DDRLError):
        K.fcnet_forward(torch.zeros(1, 11205), torch.zeros(1, 4, 19), 2)                               # CPU tensors
    with pytest.raises(DDRLError):
        K.fcnet_num_params(100, 2)                                                                     # D > 64


@pytest.mark.parametrize("scope", ["QuantrupedMultiEnv_FullyDecentral", "QuantrupedMultiEnv_Local", "QuantrupedMultiEnv_TwoSides",
                                   "QuantrupedMultiEnv_SharedDecentral", "QuantrupedMultiEnv_SingleDiagonal"])
@pytest.mark.parametrize("tvel", [False, True])
def test_obs_gather_is_bit_exact(scope, tvel):
    """a3: batched per-agent index gather == numpy fancy indexing with the reference's prefix-major index lists."""
    from ddrl_b200 import kernels as K
    from ddrl_b200.policies import ARCHITECTURES
    env = ARCHITECTURES[scope]
    table = env.gather_table(tvel)
    Ag, D = table.shape
    P = len(env.policy_names)
    k = Ag // P
    Dfull = 43 + int(tvel)
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        full = rng.standard_normal((37, Dfull)).astype(dt)
        out = K.obs_gather(_dev(full), _dev(table), P).cpu().numpy()
        ref = np.stack([np.stack([full[s][table[p * k + j]] for s in range(37) for j in range(k)]) for p in range(P)])
        assert out.shape == (P, 37 * k, D)
        assert np.array_equal(out, ref.astype(np.float32))


@pytest.mark.parametrize("scope", ["QuantrupedMultiEnv_FullyDecentral", "QuantrupedMultiEnv_TwoSides", "QuantrupedMultiEnv_Centralized"])
@pytest.mark.parametrize("mode", ["per_leg", "per_leg_norm", "global", "global_costs"])
def test_reward_split_and_action_concat_match_the_adaptor(scope, mode):
    """N2: batched per-agent reward / cost split and action concatenation == the adaptor's per-step dict loops."""
    import oracle.ddrl_oracle as O
    from ddrl_b200 import kernels as K
    from ddrl_b200.policies import ARCHITECTURES
    env = ARCHITECTURES[scope]
    Ag, A = len(env.agent_names), env.act_dim()
    rng = np.random.default_rng(Ag * 10 + A)
    S = 97
    fw = rng.standard_normal(S).astype(np.float32)
    act = (1.5 * rng.standard_normal((S, Ag, A))).astype(np.float32)           # some beyond [-1, 1]
    cfrc = 2.0 * rng.standard_normal((S, 14, 6))                               # clipped to [-1, 1] inside
    cfrc[rng.random((S, 14, 6)) < 0.5] = 0.0
    ctrl_w, con_w = 0.25, 0.025                                                # experiment-3 env_config
    cfi = env.contact_force_indices()
    rew = K.reward_split(_dev(fw), _dev(act), _dev(cfrc), _dev(env.contact_table()), ctrl_w, con_w, mode).cpu().numpy()
    ref = np.zeros((S, Ag))
    for s in range(S):
        r = O.distribute_rewards(float(fw[s]), {a: act[s, i] for i, a in enumerate(env.agent_names)}, cfrc[s], cfi,
                                 env.agent_names, ctrl_w, con_w, mode)
        ref[s] = [r[a] for a in env.agent_names]
    assert np.abs(rew - ref).max() < 1e-6 * max(1.0, np.abs(ref).max())
    table = env.action_table()
    full = K.concat_actions(_dev(act), _dev(table)).cpu().numpy()
    ai = env.action_indices()
    for s in range(0, S, 7):
        refa = O.concatenate_actions({a: np.clip(act[s, i], -1.0, 1.0) for i, a in enumerate(env.agent_names)}, ai)
        assert np.array_equal(full[s], refa.astype(np.float32))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_graph_obs_build_equals_the_reference_env(dtype):
    """N2: node features of the shared-graph env (filter-normalised index gather + ego leg quaternion) against vectors made
    by the reference's own code (tests/golden/make_graph_obs_golden.py).  float64 input: bit-exact."""
    import os
    from ddrl_b200 import kernels as K
    from ddrl_b200.policies import ARCHITECTURES
    from tests.util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "graph_obs.npz"))
    env = ARCHITECTURES["QuantrupedMultiEnv_DecentralShared_Graph"]
    obs = g["obs_full"].astype(dtype)
    table, zw = _dev(env.gather_table()), _dev(env.leg_quat_table())
    state = K.graph_obs_build(_dev(obs), table, zw, _dev(g["mean"]), _dev(g["std"]), clip=10.0).cpu().numpy()
    want = g["frozen"].astype(np.float32)
    if dtype is np.float64:
        assert np.array_equal(state, want)
    else:
        assert np.abs(state - want).max() < 2e-6 * np.abs(want).max()
    rep, node_idx = K.graph_obs_build(_dev(obs), table, zw, _dev(g["mean"]), _dev(g["std"]), clip=10.0, replicate=True)
    rep, node_idx = rep.cpu().numpy(), node_idx.cpu().numpy()
    assert np.array_equal(node_idx, np.tile(np.arange(4, dtype=np.int32), len(obs)))
    assert np.array_equal(rep.reshape(len(obs), 4, 4, 23), np.repeat(state[:, None], 4, axis=1))
    raw = K.graph_obs_build(_dev(obs), table, zw, None, None, clip=0.0).cpu().numpy()           # no filter: plain gather
    assert np.array_equal(raw[:, :, :19], obs[:, env.gather_table()].astype(np.float32))
    assert np.array_equal(raw[:, :, 19:], state[:, :, 19:])
    empty = K.graph_obs_build(_dev(obs[:0]), table, zw, _dev(g["mean"]), _dev(g["std"]))
    assert tuple(empty.shape) == (0, 4, 23)


@pytest.mark.parametrize("tag,scope", [("four", "QuantrupedMultiEnv_FullyDecentral"), ("two", "QuantrupedMultiEnv_TwoSides"),
                                       ("one", "QuantrupedMultiEnv_Centralized")])
def test_reward_split_and_action_concat_equal_reference_vectors(tag, scope):
    """N2 against vectors produced by the reference's own adaptor methods (tests/golden/make_env_glue_golden.py; its
    fw / action inputs are float32-representable).  Rewards: float64 on both sides, summed in a different order, rounded to
    float32 by the device -> 1e-6; the action scatter is bit-exact."""
    import os
    from ddrl_b200 import kernels as K
    from ddrl_b200.policies import ARCHITECTURES
    from tests.util import GOLDEN
    g = np.load(os.path.join(GOLDEN, "env_glue.npz"))
    env = ARCHITECTURES[scope]
    fw, act, cfrc = g[f"{tag}/fw"].astype(np.float32), g[f"{tag}/act"].astype(np.float32), g[f"{tag}/cfrc"]
    for mode in ("per_leg", "per_leg_norm", "global", "global_costs"):
        rew = K.reward_split(_dev(fw), _dev(act), _dev(cfrc), _dev(env.contact_table()), 0.25, 0.025, mode).cpu().numpy()
        want = g[f"{tag}/{mode}"]
        assert np.abs(rew - want).max() < 1e-6 * max(1.0, np.abs(want).max()), mode
    full = K.concat_actions(_dev(act), _dev(env.action_table())).cpu().numpy()
    assert np.array_equal(full, g[f"{tag}/actions"].astype(np.float32))
