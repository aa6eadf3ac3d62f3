"""GPU parity: GraphNet actor/critic forward + backward, GCN layer, GraphNet PPO step — vs the oracle restating
models/graph_net.py, models/gcn.py, models/shared_graphnet_glorot_uniform_init.py.
No checkpoint exists for these models: parity is against the restatement only ("unpinned")."""
import numpy as np
import pytest
import torch

from tests.util import scaled_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _O():
    import oracle.ddrl_oracle as O
    return O


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def _inputs(B, seed, adj_kind="ring"):
    O = _O()
    rng = np.random.default_rng(seed)
    state = rng.standard_normal((B, 4, 23)).astype(np.float32)
    state[..., 19:] = rng.uniform(-1, 1, size=(B, 4, 4)).astype(np.float32)   # quaternion-like leg encoding
    idx = rng.integers(0, 4, size=B).astype(np.int32)
    if adj_kind == "ring":
        adj = np.broadcast_to(O.ring_adjacency().numpy(), (B, 4, 4)).copy()
    else:  # random graphs incl. self loops, isolated receivers, weighted entries
        adj = (rng.random((B, 4, 4)) < 0.4).astype(np.float32) * rng.uniform(0.5, 2.0, size=(B, 4, 4)).astype(np.float32)
    return idx, state, adj.astype(np.float32)


def _theta(A, seed, big=False):
    O = _O()
    g = torch.Generator().manual_seed(seed)
    th = O.graphnet_wrapper_init(2 * A, g, dtype=torch.float64)
    if big:  # trained-looking magnitudes: non-zero biases, larger heads
        th = th + 0.05 * torch.randn(th.shape, generator=g, dtype=torch.float64)
    return th


@pytest.mark.parametrize("variant", [0, 1], ids=["row-per-cta", "row-per-warp"])
@pytest.mark.parametrize("B,A,adj_kind", [(1, 2, "ring"), (257, 2, "ring"), (300, 2, "random"), (64, 4, "random"),
                                          (40, 8, "ring"), (7, 2, "random"), (9, 8, "random"), (20011, 2, "ring")])
def test_graphnet_forward(B, A, adj_kind, variant):
    from ddrl_b200 import kernels as K
    O = _O()
    idx, state, adj = _inputs(B, B + A, adj_kind)
    th = _theta(A, 1, big=True)
    nan = lambda *shape: torch.full(shape, float("nan"), device="cuda")
    K.graphnet_set_variant(variant)
    try:
        lg, v = K.graphnet_forward(_dev(th.numpy(), torch.float32), _dev(idx), _dev(state), _dev(adj), A,
                                   out=(nan(B, 2 * A), nan(B)))
        torch.cuda.synchronize()
    finally:
        K.graphnet_set_variant(-1)
    th32 = th.float().double()   # the device sees float32 weights
    lg_ref, v_ref = O.graphnet_forward(th32, torch.from_numpy(idx), torch.from_numpy(state).double(),
                                       torch.from_numpy(adj).double(), 2 * A)
    assert scaled_err(lg.cpu().numpy(), lg_ref.numpy()) < TOL
    assert scaled_err(v.cpu().numpy(), v_ref.numpy()) < TOL


@pytest.mark.parametrize("B,A,adj_kind,ctas", [(1, 2, "ring", 1), (500, 2, "ring", 7), (333, 2, "random", 16),
                                               (100, 4, "random", 3)])
def test_graphnet_backward(B, A, adj_kind, ctas):
    from ddrl_b200 import kernels as K
    O = _O()
    idx, state, adj = _inputs(B, 7 * B + A, adj_kind)
    th = _theta(A, 2, big=True).float()
    rng = np.random.default_rng(3)
    dl = rng.standard_normal((B, 2 * A)).astype(np.float32)
    dv = rng.standard_normal(B).astype(np.float32)
    g = K.graphnet_backward(_dev(th.numpy()), _dev(idx), _dev(state), _dev(adj), _dev(dl), _dev(dv), A, ctas)
    torch.cuda.synchronize()
    t = th.double().requires_grad_(True)
    lg, v = O.graphnet_forward(t, torch.from_numpy(idx), torch.from_numpy(state).double(),
                               torch.from_numpy(adj).double(), 2 * A)
    (ref,) = torch.autograd.grad((lg * torch.from_numpy(dl).double()).sum() + (v * torch.from_numpy(dv).double()).sum(), t)
    g = g.cpu().numpy()
    assert scaled_err(g, ref.numpy()) < TOL
    # per variable (actor then critic), so small tensors are checked at their own scale
    o = 0
    for O_out in (2 * A, 1):
        for name, shp in O.graphnet_shapes(O_out):
            n = int(np.prod(shp))
            assert scaled_err(g[o:o + n], ref.numpy()[o:o + n]) < 2e-5, (name, O_out)
            o += n
    g2 = K.graphnet_backward(_dev(th.numpy()), _dev(idx), _dev(state), _dev(adj), _dev(dl), _dev(dv), A, ctas)
    assert torch.equal(torch.from_numpy(g), g2.cpu())   # deterministic


@pytest.mark.parametrize("F,U,bias,act", [(19, 64, True, "tanh"), (64, 64, False, "tanh"), (23, 8, True, None)])
def test_gcn_layer(F, U, bias, act):
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(F + U)
    B = 200
    x = rng.standard_normal((B, 4, F)).astype(np.float32)
    adj = np.broadcast_to(O.ring_adjacency().numpy(), (B, 4, 4)).copy()
    adj = adj * rng.uniform(0.5, 2.0, size=adj.shape).astype(np.float32)
    W = (rng.standard_normal((F, U)) / np.sqrt(F)).astype(np.float32)
    b = rng.standard_normal(U).astype(np.float32) if bias else None
    y = K.gcn_forward(_dev(x), _dev(adj), _dev(W), _dev(b) if bias else None, act)
    ref = O.gcn_layer(torch.from_numpy(x).double(), torch.from_numpy(adj).double(), torch.from_numpy(W).double(),
                      torch.from_numpy(b).double() if bias else None, act)
    assert scaled_err(y.cpu().numpy(), ref.numpy()) < TOL


def test_graphnet_ppo_iteration_runs_and_first_step_matches_oracle():
    """Shared GraphNet policy: forward/sample/GAE/standardise, then ONE minibatch step compared with the oracle
    (autograd float64 through the restated GraphNet + PPO loss + clip + TF1 Adam)."""
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import GraphNetLearner
    O = _O()
    A, T, N = 2, 8, 16
    C = N * 4
    R = T * C
    cfg = PPOConfig(num_sgd_iter=1, sgd_minibatch_size=R)
    cfg_o = O.PPOConfig(num_sgd_iter=1, sgd_minibatch_size=R)
    idx, state, adj = _inputs(R, 21, "ring")
    idx = np.tile(np.arange(4, dtype=np.int32), R // 4)
    bidx, bstate, badj = _inputs(C, 22, "ring")
    rng = np.random.default_rng(23)
    rewards = (0.3 + 0.5 * rng.standard_normal((T, C))).astype(np.float32)
    dones = (rng.random((T, N)) < 0.05).astype(np.uint8)
    eps = rng.standard_normal((T, C, A)).astype(np.float32)
    th0 = _theta(A, 5, big=True).float()
    L = GraphNetLearner(A, cfg, "cuda", theta=th0.reshape(1, -1))
    perms = np.zeros((1, 1), np.int32)
    stats = L.learn_on_rollout(_dev(idx.reshape(T, C)), _dev(state.reshape(T, C, 4, 23)), _dev(adj.reshape(T, C, 4, 4)),
                               _dev(bidx), _dev(bstate), _dev(badj), _dev(rewards), _dev(dones), _dev(eps), _dev(perms))
    torch.cuda.synchronize()
    # oracle ------------------------------------------------------------------------------------------------------
    t = th0.double()
    fwd = lambda th, obs: O.graphnet_forward(th, obs[0], obs[1], obs[2], 2 * A)
    ti, ts, ta = torch.from_numpy(idx), torch.from_numpy(state).double(), torch.from_numpy(adj).double()
    with torch.no_grad():
        lg, v = fwd(t, (ti, ts, ta))
        act = O.dg_sample(lg, torch.from_numpy(eps.reshape(R, A)).double())
        logp = O.dg_logp(lg, act)
        _, vb = fwd(t, (torch.from_numpy(bidx), torch.from_numpy(bstate).double(), torch.from_numpy(badj).double()))
    adv, vt = O.gae_recurrence(rewards, v.numpy().astype(np.float32).reshape(T, C), np.repeat(dones, 4, axis=1),
                               vb.numpy().astype(np.float32), cfg.gamma, cfg.lambda_)
    adv_s = O.standardized(adv.reshape(-1))

    class _Obs:  # row-sliceable tuple observation
        def __getitem__(self, rows):
            return (ti[rows], ts[rows], ta[rows])
    batch = {"obs": _Obs(), "actions": act, "old_logits": lg, "old_logp": logp, "vf_preds": v,
             "advantages": torch.from_numpy(adv_s).double(), "value_targets": torch.from_numpy(vt.reshape(-1)).double()}
    st = O.AdamState.zeros(t.numel(), torch.float64, cfg_o)
    new_t, s_ref, g_ref, gn_ref = O.sgd_minibatch_step(t, st, fwd, batch, slice(0, R), cfg_o.kl_coeff, cfg_o)
    assert scaled_err(L.grad.cpu().numpy().reshape(-1), g_ref.numpy()) < 2e-5
    assert abs(float(L.gnorm[0]) - gn_ref) < 1e-5 * gn_ref
    upd_d = L.theta.cpu().numpy().reshape(-1).astype(np.float64) - th0.numpy()
    assert scaled_err(L.theta.cpu().numpy().reshape(-1), new_t.numpy()) < TOL
    for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "vf_explained_var"):
        assert abs(stats[0][k] - s_ref[k]) < 1e-4 * max(1.0, abs(s_ref[k])), (k, stats[0][k], s_ref[k])


# ---- the graph-layer variants the reference carries but does not wire into a model ---------------------
def _rand_adj(rng, B, selfloops=False):
    adj = (rng.random((B, 4, 4)) < 0.5).astype(np.float32)
    ring = _O().ring_adjacency().numpy()
    adj[: B // 2] = ring                      # half the batch: the quantruped ring
    if not selfloops:
        adj[:, np.arange(4), np.arange(4)] = 0.0
    return adj


@pytest.mark.parametrize("F,U,bias,act", [(23, 64, False, "tanh"), (64, 64, True, "tanh"), (19, 32, True, None)])
def test_mpnn2_layer(F, U, bias, act):
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(F + U)
    B = 257
    x = rng.standard_normal((B, 4, F)).astype(np.float32)
    adj = _rand_adj(rng, B)
    Wm = (rng.standard_normal((2 * F, U)) * 0.2).astype(np.float32)
    Wu = (rng.standard_normal((F + U, U)) * 0.2).astype(np.float32)
    b = rng.standard_normal(U).astype(np.float32)
    y = K.mpnn2_forward(_dev(x), _dev(adj), _dev(Wm), _dev(Wu), _dev(b) if bias else None, act)
    d = lambda a: torch.from_numpy(a).double()
    ref = O.mpnn2_layer(d(x), d(adj), d(Wm), d(Wu), d(b) if bias else None, act)
    assert scaled_err(y.cpu().numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize("F,U,bias,act", [(23, 64, False, "tanh"), (64, 48, True, None)])
def test_gat1_layer(F, U, bias, act):
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(F * U)
    B = 130
    x = rng.standard_normal((B, 4, F)).astype(np.float32)
    adj = _rand_adj(rng, B, selfloops=True)
    Wp = (rng.standard_normal((F, U)) * 0.2).astype(np.float32)
    wa = (rng.standard_normal((2 * U, 1)) * 0.3).astype(np.float32)
    b = rng.standard_normal(U).astype(np.float32)
    y = K.gat1_forward(_dev(x), _dev(adj), _dev(Wp), _dev(wa), _dev(b) if bias else None, act)
    d = lambda a: torch.from_numpy(a).double()
    ref = O.gat1_layer(d(x), d(adj), d(Wp), d(wa), d(b) if bias else None, act)
    assert scaled_err(y.cpu().numpy(), ref.numpy()) < 1e-5


def test_symm_norm_and_segment_softmax():
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(3)
    adj = (rng.random((64, 4, 4)) < 0.6).astype(np.float32)
    adj[:, :, 0] = 1.0                         # every row has a neighbour (zero degree is NaN in the reference too)
    out = K.symm_norm(_dev(adj)).cpu().numpy()
    ref = O.symm_norm(torch.from_numpy(adj).double()).numpy()
    assert scaled_err(out, ref) < 1e-6
    z = np.zeros((1, 4, 4), np.float32)        # zero-degree rows: NaN, like tf (0 ** -0.5 = inf, inf * 0 = nan)
    assert np.isnan(K.symm_norm(_dev(z)).cpu().numpy()).all()
    E, S = 1000, 37
    data = rng.standard_normal((E, 1)).astype(np.float32)
    seg = rng.integers(0, S, size=E).astype(np.int32)
    sm = K.segment_softmax(_dev(data), _dev(seg), S).cpu().numpy()
    ref = O.segment_softmax(torch.from_numpy(data).double(), torch.from_numpy(seg).long(), S).numpy()
    assert scaled_err(sm, ref) < 1e-6
    sums = np.zeros(S)
    np.add.at(sums, seg, sm[:, 0])
    assert np.allclose(sums[np.bincount(seg, minlength=S) > 0], 1.0, atol=1e-5)      # every non-empty segment sums to 1
    from ddrl_b200._lib import DDRLError
    with pytest.raises(DDRLError):
        K.segment_softmax(_dev(data), _dev(seg), S - 5)


@pytest.mark.parametrize("F,U,bias,act,ctas", [(19, 64, True, "tanh", None), (64, 64, False, None, 7), (23, 32, True, "tanh", 300)])
def test_mpnn2_backward_matches_autograd(F, U, bias, act, ctas):
    """Backward of MPNN2: gradients w.r.t. the inputs, W_msg, W_upd and the bias against the oracle's
    float64 autograd through `mpnn2_layer` (random graphs incl. isolated receivers); more CTAs than samples allowed."""
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(3 * F + U)
    B = 257
    x = rng.standard_normal((B, 4, F)).astype(np.float32)
    adj = _rand_adj(rng, B)
    Wm = (rng.standard_normal((2 * F, U)) * 0.2).astype(np.float32)
    Wu = (rng.standard_normal((F + U, U)) * 0.2).astype(np.float32)
    b = rng.standard_normal(U).astype(np.float32)
    dy = rng.standard_normal((B, 4, U)).astype(np.float32)
    dx, gWm, gWu, gb = K.mpnn2_backward(_dev(x), _dev(adj), _dev(Wm), _dev(Wu), _dev(b) if bias else None, _dev(dy), act, ctas)
    d = lambda a: torch.from_numpy(a).double().requires_grad_(True)
    tx, tWm, tWu, tb = d(x), d(Wm), d(Wu), d(b)
    ref = O.mpnn2_layer(tx, torch.from_numpy(adj).double(), tWm, tWu, tb if bias else None, act)
    (ref * torch.from_numpy(dy).double()).sum().backward()
    assert scaled_err(dx.cpu().numpy(), tx.grad.numpy()) < 1e-5
    assert scaled_err(gWm.cpu().numpy(), tWm.grad.numpy()) < 1e-5
    assert scaled_err(gWu.cpu().numpy(), tWu.grad.numpy()) < 1e-5
    if bias:
        assert scaled_err(gb.cpu().numpy(), tb.grad.numpy()) < 1e-5


@pytest.mark.parametrize("F,U,bias,act,ctas", [(23, 64, False, "tanh", None), (64, 48, True, None, 5), (19, 64, True, "tanh", 200)])
def test_gat1_backward_matches_autograd(F, U, bias, act, ctas):
    """Backward of GAT1 (self loops, leaky-relu attention logits, softmax over the senders of each receiver): gradients
    w.r.t. the inputs, W_pre, w_att and the bias against the oracle's float64 autograd through `gat1_layer`."""
    from ddrl_b200 import kernels as K
    O = _O()
    rng = np.random.default_rng(F * U + 1)
    B = 130
    x = rng.standard_normal((B, 4, F)).astype(np.float32)
    adj = _rand_adj(rng, B, selfloops=True)
    Wp = (rng.standard_normal((F, U)) * 0.2).astype(np.float32)
    wa = (rng.standard_normal((2 * U, 1)) * 0.3).astype(np.float32)
    b = rng.standard_normal(U).astype(np.float32)
    dy = rng.standard_normal((B, 4, U)).astype(np.float32)
    dx, gWp, gwa, gb = K.gat1_backward(_dev(x), _dev(adj), _dev(Wp), _dev(wa), _dev(b) if bias else None, _dev(dy), act, ctas)
    d = lambda a: torch.from_numpy(a).double().requires_grad_(True)
    tx, tWp, twa, tb = d(x), d(Wp), d(wa), d(b)
    ref = O.gat1_layer(tx, torch.from_numpy(adj).double(), tWp, twa, tb if bias else None, act)
    (ref * torch.from_numpy(dy).double()).sum().backward()
    assert scaled_err(dx.cpu().numpy(), tx.grad.numpy()) < 1e-5
    assert scaled_err(gWp.cpu().numpy(), tWp.grad.numpy()) < 1e-5
    assert scaled_err(gwa.cpu().numpy(), twa.grad.numpy().reshape(-1)) < 1e-5
    if bias:
        assert scaled_err(gb.cpu().numpy(), tb.grad.numpy()) < 1e-5
