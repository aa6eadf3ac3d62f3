"""GPU: the RLlib-facing learner surface — `FCNetLearner.learn_on_batch` on postprocessed sample-batch
columns and `rllib_policy.PPOPolicyGroup` (compute_actions / postprocess_fragments / learn_on_batch / weights) against
the oracle's restatement of the same RLlib steps."""
import numpy as np
import pytest
import torch

from tests.util import ARCHS, ckpt_theta, scaled_err, synth_obs

pytestmark = pytest.mark.gpu

ARCH, SCOPE = "FullyDecentral", "QuantrupedMultiEnv_FullyDecentral"
COLS = ("obs", "act", "logits", "logp", "value", "adv", "vtarg")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_learn_on_batch_is_the_sgd_phase_of_learn_on_rollout(mode):
    """Feeding the columns a rollout iteration prepared (filtered obs, sampled actions, old logits / logp, value
    predictions, standardised advantages, value targets) to `learn_on_batch` reproduces that iteration bit for bit."""
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import FCNetLearner
    theta0, filt, D, A = ckpt_theta(ARCH)
    P, T, C, E, NB = theta0.shape[0], 16, 32, 2, 4
    R = T * C
    cfg = PPOConfig(num_sgd_iter=E, sgd_minibatch_size=R // NB)
    rng = np.random.default_rng(21)
    raw, boot = synth_obs(filt, R, 1).reshape(P, T, C, D), synth_obs(filt, C, 2)
    rewards = (0.3 + 0.5 * rng.standard_normal((P, T, C))).astype(np.float32)
    dones = (rng.random((T, C)) < 0.05).astype(np.uint8)
    eps = rng.standard_normal((P, T, C, A)).astype(np.float32)
    perms = np.stack([np.stack([rng.permutation(NB) for _ in range(E)]) for _ in range(P)]).astype(np.int32)
    shuffle = np.stack([rng.permutation(R) for _ in range(P)]).astype(np.int32)

    def learner():
        # (fixed-order reduction of the partial gradients: the two learners are compared bit for bit)
        L = FCNetLearner(P, D, A, cfg, "cuda", theta=torch.from_numpy(theta0), mode=mode, atomic_reduce=False)
        L.filt_n.copy_(torch.tensor([f[0] for f in filt]))
        L.filt_M.copy_(torch.from_numpy(np.stack([f[1] for f in filt])))
        L.filt_S.copy_(torch.from_numpy(np.stack([f[2] for f in filt])))
        return L

    La = learner()
    stats_a = La.learn_on_rollout(_dev(raw), _dev(boot), _dev(rewards), _dev(dones), _dev(eps), _dev(perms), _dev(shuffle))
    torch.cuda.synchronize()
    cols = [La._bufs[n].clone() for n in COLS]
    Lb = learner()
    stats_b = Lb.learn_on_batch(*cols, _dev(perms), _dev(shuffle), standardize=False)
    torch.cuda.synchronize()
    assert not torch.equal(La.theta, torch.from_numpy(theta0).cuda())
    assert torch.equal(La.theta, Lb.theta) and torch.equal(La.m, Lb.m) and torch.equal(La.v, Lb.v)
    assert torch.equal(La.beta_pow, Lb.beta_pow) and np.array_equal(La.kl_coeff_host, Lb.kl_coeff_host)
    assert stats_a == stats_b
    # standardising an already standardised column changes it only at round-off level
    Lc = learner()
    stats_c = Lc.learn_on_batch(*cols, _dev(perms), _dev(shuffle), standardize=True)
    for p in range(P):
        for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy"):
            assert abs(stats_c[p][k] - stats_a[p][k]) < 1e-4 * max(1.0, abs(stats_a[p][k])), (k, stats_c[p][k], stats_a[p][k])
    with pytest.raises(Exception, match="float32 CUDA tensor"):
        Lc.learn_on_batch(cols[0], cols[1], cols[2], cols[3], cols[4], cols[5][:, :-1], cols[6], _dev(perms))


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_policy_group_serves_the_rllib_calls(mode):
    import oracle.ddrl_oracle as O
    from ddrl_b200 import rllib_policy as RP
    from ddrl_b200.checkpoint import theta_to_variables
    theta0, _, D, A = ckpt_theta(ARCH)
    names = ARCHS[ARCH][0]
    P, T, C, E, MB, SEED = theta0.shape[0], 8, 65, 2, 128, 3
    R = T * C                                              # 520 rows: 4 whole minibatches of 128 + a ragged tail of 8
    g = RP.PPOPolicyGroup(SCOPE, {"num_sgd_iter": E, "sgd_minibatch_size": MB, "seed": SEED}, mode=mode)
    assert g.policy_names == list(names) and (g.D, g.A) == (D, A)
    g.set_weights({pid: theta_to_variables(pid, theta0[i], D, A) for i, pid in enumerate(names)})
    w = g.get_weights()
    assert list(w[names[0]])[0] == f"{names[0]}/fc_1/kernel" and w[names[0]][f"{names[0]}/fc_1/kernel"].shape == (D, 64)
    assert np.array_equal(g.learner.theta.cpu().numpy(), theta0)

    rng = np.random.default_rng(4)
    obs = rng.standard_normal((P, R, D)).astype(np.float32)          # what the workers' MeanStdFilter would hand over
    boot = rng.standard_normal((P, C, D)).astype(np.float32)
    rewards = (0.3 + 0.5 * rng.standard_normal((P, T, C))).astype(np.float32)
    dones = (rng.random((T, C)) < 0.1).astype(np.uint8)

    # ---- compute_actions -------------------------------------------------------------------------------------------
    out = g.compute_actions({pid: obs[i] for i, pid in enumerate(names)})
    last = g.compute_actions({pid: boot[i] for i, pid in enumerate(names)}, explore=False)
    replay = np.random.RandomState(SEED)
    noise = replay.standard_normal((P, R, A)).astype(np.float32)
    for i, pid in enumerate(names):
        act, state, info = out[pid]
        assert state == [] and act.shape == (R, A) and set(info) == {RP.ACTION_DIST_INPUTS, RP.ACTION_LOGP, RP.VF_PREDS}
        logits, value = O.fcnet_forward(torch.from_numpy(theta0[i]).double(), torch.from_numpy(obs[i]).double(), 2 * A)
        want_act = O.dg_sample(logits, torch.from_numpy(noise[i]).double())
        assert scaled_err(info[RP.ACTION_DIST_INPUTS], logits.numpy()) < 1e-5
        assert scaled_err(info[RP.VF_PREDS], value.numpy()) < 1e-5
        assert scaled_err(act, want_act.numpy()) < 1e-5
        assert scaled_err(info[RP.ACTION_LOGP], O.dg_logp(logits, want_act).numpy()) < 1e-5
        mean = last[pid][2][RP.ACTION_DIST_INPUTS][:, :A]
        assert np.array_equal(last[pid][0], mean)                     # explore=False: the distribution mean

    # ---- postprocess_fragments (postprocess_ppo_gae, batched) ------------------------------------------------------------
    vf = np.stack([out[pid][2][RP.VF_PREDS] for pid in names]).reshape(P, T, C)
    last_v = np.stack([last[pid][2][RP.VF_PREDS] for pid in names])
    adv, vt = g.postprocess_fragments(rewards, vf, dones, last_v)
    for i in range(P):
        a_ref, v_ref = O.gae_recurrence(rewards[i], vf[i], dones, last_v[i], 0.99, 0.95)
        assert scaled_err(adv[i], a_ref) < 1e-5 and scaled_err(vt[i], v_ref) < 1e-5

    # ---- learn_on_batch ----------------------------------------------------------------------------------------------
    batches = {pid: {RP.OBS: obs[i], RP.ACTIONS: out[pid][0], RP.ACTION_DIST_INPUTS: out[pid][2][RP.ACTION_DIST_INPUTS],
                     RP.ACTION_LOGP: out[pid][2][RP.ACTION_LOGP], RP.VF_PREDS: out[pid][2][RP.VF_PREDS],
                     RP.ADVANTAGES: adv[i].reshape(-1), RP.VALUE_TARGETS: vt[i].reshape(-1)} for i, pid in enumerate(names)}
    res = g.learn_on_batch(batches)
    assert g.num_steps_trained == 512 and set(res) == set(names)
    shuffle, perms = RP.draw_minibatch_order(replay, P, R, E, 4)        # the draws the group made after the action noise
    cfg_o = O.PPOConfig(num_sgd_iter=E, sgd_minibatch_size=MB)
    fwd = lambda th, xx: O.fcnet_forward(th, xx, 2 * A)
    for i, pid in enumerate(names):
        b = batches[pid]
        idx = shuffle[i][:512].astype(np.int64)
        adv_std = O.standardized(np.asarray(b[RP.ADVANTAGES], dtype=np.float32))          # over all 520 rows
        thetas = {}
        for dt in (torch.float64, torch.float32):
            t = lambda a: torch.from_numpy(np.asarray(a)[idx]).to(dt)
            ob = {"obs": t(b[RP.OBS]), "actions": t(b[RP.ACTIONS]), "old_logits": t(b[RP.ACTION_DIST_INPUTS]),
                  "old_logp": t(b[RP.ACTION_LOGP]), "vf_preds": t(b[RP.VF_PREDS]), "advantages": t(adv_std),
                  "value_targets": t(b[RP.VALUE_TARGETS])}
            thetas[dt], st = O.sgd_loop(torch.from_numpy(theta0[i]).to(dt), O.AdamState.zeros(theta0.shape[1], dt, cfg_o), fwd,
                                        ob, perms[i], cfg_o.kl_coeff, cfg_o)
            if dt is torch.float64:
                ref_stats = st
        got = g.learner.theta[i].cpu().numpy().astype(np.float64)
        upd = thetas[torch.float64].numpy() - theta0[i]
        err_dev = scaled_err(got - theta0[i], upd)
        err_twin = scaled_err(thetas[torch.float32].numpy().astype(np.float64) - theta0[i], upd)
        assert err_dev < 10.0 * err_twin + 1e-5, (err_dev, err_twin)
        ls = res[pid]["learner_stats"]
        for k in ("total_loss", "policy_loss", "vf_loss", "kl", "entropy", "vf_explained_var"):
            assert abs(ls[k] - ref_stats[k]) < 1e-4 * max(1.0, abs(ref_stats[k])), (k, ls[k], ref_stats[k])
        assert {"cur_kl_coeff", "cur_lr"} <= set(ls)
    moved = g.get_weights()
    assert not np.array_equal(moved[names[0]][f"{names[0]}/fc_out/kernel"], w[names[0]][f"{names[0]}/fc_out/kernel"])

    # ---- errors ----------------------------------------------------------------------------------------------------------
    bad = {pid: dict(v) for pid, v in batches.items()}
    bad[names[1]][RP.VF_PREDS] = bad[names[1]][RP.VF_PREDS][:100]
    with pytest.raises(RP.BatchError, match="equally long"):
        g.learn_on_batch(bad)
    with pytest.raises(NotImplementedError):
        RP.PPOPolicyGroup("QuantrupedMultiEnv_DecentralShared_Graph")
