"""GPU: the two-launch GraphNet SGD step (`ddrl_graphnet_train_step`: warp-per-row forward + PPO loss + backward to
the layer inputs, then thread-owned weight-gradient accumulation) against the validated three-kernel step
(`ddrl_graphnet_forward` + `ddrl_ppo_loss_grad` + `ddrl_graphnet_backward`) and the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_graphnet import _O, _dev, _inputs, _theta
from tests.util import scaled_err

pytestmark = pytest.mark.gpu


def _batch(B, A, seed):
    O = _O()
    rng = np.random.default_rng(seed)
    old_logits = (0.3 * rng.standard_normal((B, 2 * A))).astype(np.float32)
    actions = (old_logits[:, :A] + np.exp(old_logits[:, A:]) * rng.standard_normal((B, A))).astype(np.float32)
    old_logp = O.dg_logp(torch.from_numpy(old_logits), torch.from_numpy(actions)).numpy().astype(np.float32)
    vf_preds = rng.standard_normal(B).astype(np.float32)
    adv = rng.standard_normal(B).astype(np.float32)
    vtarg = (vf_preds + 3.0 * rng.standard_normal(B)).astype(np.float32)        # some rows beyond vf_clip_param
    return actions, old_logits, old_logp, vf_preds, adv, vtarg


@pytest.mark.parametrize("B,A,adj_kind,ctas", [(1, 2, "ring", 1), (300, 2, "ring", 7), (333, 2, "random", 16),
                                               (100, 4, "random", 3), (57, 8, "ring", 5), (4099, 2, "ring", 74)])
def test_two_launch_step_equals_three_kernel_step(B, A, adj_kind, ctas):
    from ddrl_b200 import kernels as K
    from ddrl_b200._lib import PPOHyper
    idx, state, adj = _inputs(B, 11 * B + A, adj_kind)
    th = _dev(_theta(A, 3, big=True).float().numpy())
    NP = th.numel()
    cols = [_dev(c) for c in _batch(B, A, B + 1)]
    act, ol, olp, vfp, adv, vt = cols
    d_idx, d_st, d_adj = _dev(idx), _dev(state), _dev(adj)
    kl = torch.tensor([0.3], dtype=torch.float32, device="cuda")
    hyper = PPOHyper(0.2, 2.0, 0.5, 0.01, 1.0 / B)
    # validated path ---------------------------------------------------------------------------------------------------
    lg, v = K.graphnet_forward(th, d_idx, d_st, d_adj, A)
    LG = 3
    dl, dv = torch.empty(B, 2 * A, device="cuda"), torch.empty(B, device="cuda")
    sp_ref = torch.empty(1, LG, K.NSTAT, dtype=torch.float64, device="cuda")
    K.ppo_loss_grad(lg.reshape(1, B, 2 * A), v.reshape(1, B), act, ol, olp, vfp, adv, vt, A, kl, hyper, LG, dl, dv, sp_ref)
    g_ref = K.graphnet_backward(th, d_idx, d_st, d_adj, dl, dv, A, ctas)
    # two-launch path --------------------------------------------------------------------------------------------------
    parts = K.graphnet_train_stat_parts(B)
    gpart = torch.full((ctas, K.part_stride(NP)), float("nan"), device="cuda")
    spart = torch.full((parts, K.NSTAT), float("nan"), dtype=torch.float64, device="cuda")
    K.graphnet_train_step(th, d_idx, d_st, d_adj, act, ol, olp, vfp, adv, vt, A, kl, hyper, ctas, gpart, spart)
    torch.cuda.synchronize()
    g_new = gpart[:, :NP].sum(dim=0)
    assert torch.isfinite(g_new).all() and torch.isfinite(spart).all()
    assert scaled_err(g_new.cpu().numpy(), g_ref.cpu().numpy()) < 1e-5
    o = 0
    O = _O()
    for O_out in (2 * A, 1):
        for name, shp in O.graphnet_shapes(O_out):
            n = int(np.prod(shp))
            assert scaled_err(g_new[o:o + n].cpu().numpy(), g_ref[o:o + n].cpu().numpy()) < 2e-5, (name, O_out)
            o += n
    s_new, s_ref = spart.sum(dim=0).cpu().numpy(), sp_ref.sum(dim=(0, 1)).cpu().numpy()
    assert np.allclose(s_new, s_ref, rtol=1e-6, atol=1e-9), (s_new, s_ref)


def test_learner_iteration_with_the_two_launch_step():
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import GraphNetLearner
    A, T, N = 2, 8, 16
    C, R = N * 4, 8 * 16 * 4
    cfg = PPOConfig(num_sgd_iter=2, sgd_minibatch_size=R // 4)
    idx, state, adj = _inputs(R, 21, "ring")
    idx = np.tile(np.arange(4, dtype=np.int32), R // 4)
    bidx, bstate, badj = _inputs(C, 22, "ring")
    rng = np.random.default_rng(23)
    rewards = (0.3 + 0.5 * rng.standard_normal((T, C))).astype(np.float32)
    dones = (rng.random((T, N)) < 0.05).astype(np.uint8)
    eps = rng.standard_normal((T, C, A)).astype(np.float32)
    perms = np.stack([rng.permutation(4) for _ in range(2)]).astype(np.int32)
    shuffle = rng.permutation(R).astype(np.int32)
    th0 = _theta(A, 5, big=True).float()
    res = []
    for kind in ("three-kernel", "two-launch"):
        L = GraphNetLearner(A, cfg, "cuda", theta=th0.reshape(1, -1), step=kind)
        stats = L.learn_on_rollout(_dev(idx.reshape(T, C)), _dev(state.reshape(T, C, 4, 23)), _dev(adj.reshape(T, C, 4, 4)),
                                   _dev(bidx), _dev(bstate), _dev(badj), _dev(rewards), _dev(dones), _dev(eps), _dev(perms),
                                   _dev(shuffle))
        torch.cuda.synchronize()
        res.append((L.theta.cpu().numpy().reshape(-1), stats[0]))
    (th_a, st_a), (th_b, st_b) = res
    upd = th_a.astype(np.float64) - th0.numpy()
    assert scaled_err(th_b.astype(np.float64) - th0.numpy(), upd) < 0.05       # FP32 order differs; Adam amplifies
    for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "vf_explained_var"):
        assert abs(st_a[k] - st_b[k]) < 1e-4 * max(1.0, abs(st_a[k])), (k, st_a[k], st_b[k])


# ---- the tensor-core step (csrc/graphnet_tc.cu): FMA/MUFU hyper-encoder + tcgen05 MPNN / head / weight-gradient GEMMs ----------
TC_GRAD_TOL = 5e-5      # stated tolerance of the fp16 hi/lo split GEMMs (as for the FCNet tensor-core step); forward stats 1e-5


@pytest.mark.parametrize("B,A,adj_kind,ctas", [(1, 2, "ring", 1), (300, 2, "ring", 7), (333, 2, "random", 16),
                                               (100, 4, "random", 3), (57, 8, "ring", 5), (4099, 2, "ring", 74),
                                               (640, 2, "selfloops", 3)])
def test_tc_step_equals_three_kernel_step(B, A, adj_kind, ctas):
    """`ddrl_graphnet_train_step_tc` without the fused tail: per-variable gradients and loss statistics of one minibatch
    against the validated FP32 path (forward + ppo_loss_grad + backward) — ring, random and self-loop graphs (up to 4 encoded
    nodes per row), ragged tiles, CTAs without rows."""
    from ddrl_b200 import kernels as K
    from ddrl_b200._lib import PPOHyper
    if adj_kind == "selfloops":
        idx, state, adj = _inputs(B, 11 * B + A, "random")
        rng = np.random.default_rng(B)
        adj[:, np.arange(4), np.arange(4)] = (rng.random((B, 4)) < 0.5).astype(np.float32)
    else:
        idx, state, adj = _inputs(B, 11 * B + A, adj_kind)
    th = _dev(_theta(A, 3, big=True).float().numpy())
    NP = th.numel()
    act, ol, olp, vfp, adv, vt = [_dev(c) for c in _batch(B, A, B + 1)]
    d_idx, d_st, d_adj = _dev(idx), _dev(state), _dev(adj)
    kl = torch.tensor([0.3], dtype=torch.float32, device="cuda")
    hyper = PPOHyper(0.2, 2.0, 0.5, 0.01, 1.0 / B)
    lg, v = K.graphnet_forward(th, d_idx, d_st, d_adj, A)
    LG = 3
    dl, dv = torch.empty(B, 2 * A, device="cuda"), torch.empty(B, device="cuda")
    sp_ref = torch.empty(1, LG, K.NSTAT, dtype=torch.float64, device="cuda")
    K.ppo_loss_grad(lg.reshape(1, B, 2 * A), v.reshape(1, B), act, ol, olp, vfp, adv, vt, A, kl, hyper, LG, dl, dv, sp_ref)
    g_ref = K.graphnet_backward(th, d_idx, d_st, d_adj, dl, dv, A, max(1, min(ctas, B)))
    gpart = torch.full((ctas, K.part_stride(NP)), float("nan"), device="cuda")
    spart = torch.zeros(2 * ctas, K.NSTAT, dtype=torch.float64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    K.graphnet_train_step_tc(th, d_idx, d_st, d_adj, act, ol, olp, vfp, adv, vt, A, B, None, None, kl, hyper, ctas, gpart, spart,
                             status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    g_new = gpart[:, :NP].sum(dim=0)
    assert torch.isfinite(g_new).all() and torch.isfinite(spart).all()
    o = 0
    O = _O()
    for O_out in (2 * A, 1):
        for name, shp in O.graphnet_shapes(O_out):
            n = int(np.prod(shp))
            err = scaled_err(g_new[o:o + n].cpu().numpy(), g_ref[o:o + n].cpu().numpy())
            assert err < TC_GRAD_TOL, (name, O_out, err)
            o += n
    s_new, s_ref = spart.sum(dim=0).cpu().numpy(), sp_ref.sum(dim=(0, 1)).cpu().numpy()
    assert np.allclose(s_new, s_ref, rtol=1e-5, atol=1e-7 * B), (s_new, s_ref)


@pytest.mark.parametrize("kind", ["tc", "three-kernel"])
def test_graphnet_learner_all_steps_and_final_weights_match_oracle(kind):
    """The shared GraphNet policy through a whole learner iteration — forward / sample / GAE / standardise / shuffle, then
    2 epochs x 4 minibatch steps (clip + TF1 Adam) — against the oracle's float64 trajectory: per-step loss statistics of the
    last epoch and the FINAL weights (not only the first step), for the tensor-core step and the FP32 step."""
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import GraphNetLearner
    O = _O()
    A, T, N, E, NB = 2, 8, 24, 2, 4
    C, R = N * 4, 8 * 24 * 4
    cfgd = dict(num_sgd_iter=E, sgd_minibatch_size=R // NB)
    cfg, cfg_o = PPOConfig(**cfgd), O.PPOConfig(**cfgd)
    idx, state, adj = _inputs(R, 31, "ring")
    idx = np.tile(np.arange(4, dtype=np.int32), R // 4)
    bidx, bstate, badj = _inputs(C, 32, "ring")
    rng = np.random.default_rng(33)
    rewards = (0.3 + 0.5 * rng.standard_normal((T, C))).astype(np.float32)
    dones = (rng.random((T, N)) < 0.05).astype(np.uint8)
    eps = rng.standard_normal((T, C, A)).astype(np.float32)
    perms = np.stack([rng.permutation(NB) for _ in range(E)]).astype(np.int32)
    shuffle = rng.permutation(R).astype(np.int32)
    th0 = _theta(A, 5, big=True).float()
    L = GraphNetLearner(A, cfg, "cuda", theta=th0.reshape(1, -1), step=kind)
    stats = L.learn_on_rollout(_dev(idx.reshape(T, C)), _dev(state.reshape(T, C, 4, 23)), _dev(adj.reshape(T, C, 4, 4)),
                               _dev(bidx), _dev(bstate), _dev(badj), _dev(rewards), _dev(dones), _dev(eps), _dev(perms),
                               _dev(shuffle))
    torch.cuda.synchronize()

    def oracle(dtype):
        t = th0.to(dtype)
        fwd = lambda th, obs: O.graphnet_forward(th, obs[0], obs[1], obs[2], 2 * A)
        ti, ts_, ta = torch.from_numpy(idx), torch.from_numpy(state).to(dtype), torch.from_numpy(adj).to(dtype)
        with torch.no_grad():
            lg, v = fwd(t, (ti, ts_, ta))
            act = O.dg_sample(lg, torch.from_numpy(eps.reshape(R, A)).to(dtype))
            logp = O.dg_logp(lg, act)
            _, vb = fwd(t, (torch.from_numpy(bidx), torch.from_numpy(bstate).to(dtype), torch.from_numpy(badj).to(dtype)))
        adv, vt = O.gae_recurrence(rewards, v.numpy().astype(np.float32).reshape(T, C), np.repeat(dones, 4, axis=1),
                                   vb.numpy().astype(np.float32), cfg.gamma, cfg.lambda_)
        adv_s = O.standardized(adv.reshape(-1))
        sh = torch.from_numpy(shuffle.astype(np.int64))

        class _Obs:  # row-sliceable tuple observation (already shuffled)
            def __getitem__(self, rows):
                return (ti[sh][rows], ts_[sh][rows], ta[sh][rows])
        batch = {"obs": _Obs(), "actions": act[sh], "old_logits": lg[sh], "old_logp": logp[sh], "vf_preds": v[sh],
                 "advantages": torch.from_numpy(adv_s).to(dtype)[sh], "value_targets": torch.from_numpy(vt.reshape(-1)).to(dtype)[sh]}
        st = O.AdamState.zeros(t.numel(), dtype, cfg_o)
        return O.sgd_loop(t, st, fwd, batch, perms, cfg_o.kl_coeff, cfg_o)
    th64, s64 = oracle(torch.float64)
    th32, _ = oracle(torch.float32)
    upd_o = th64.numpy() - th0.numpy().astype(np.float64)
    th_d = L.theta.cpu().numpy().reshape(-1)
    err_dev = scaled_err(th_d.astype(np.float64) - th0.numpy(), upd_o)
    err_twin = scaled_err(th32.numpy().astype(np.float64) - th0.numpy(), upd_o)
    assert err_dev < 10.0 * err_twin + 1e-5, (err_dev, err_twin)      # no further from float64 than 10x the float32 twin
    assert scaled_err(th_d, th64.numpy()) < 2e-3
    for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "vf_explained_var"):
        assert abs(stats[0][k] - s64[k]) < 1e-4 * max(1.0, abs(s64[k])), (k, stats[0][k], s64[k])
