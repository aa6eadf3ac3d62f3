"""Host-logic dry run of `GraphNetLearner` (CPU, own process — it monkeypatches torch; run by tests/test_host.py).

Every kernel wrapper the learner touches (GraphNet forward / backward, DiagGaussian sample, GAE, StandardizeFields, PPO loss
gradient, partial reduction, clip + TF1 Adam) is replaced by the ORACLE's float64 restatement, "cuda" is mapped to "cpu", and
two GPU tests are executed unchanged: one learner iteration against the oracle (tests/test_gpu_graphnet.py) and
`learn_on_batch` == the SGD phase of `learn_on_rollout` (tests/test_gpu_graphnet_batch.py).  It checks the orchestration —
row order, slicing, shuffle, which statistics go where, KL update — and no kernel.  Test infrastructure only."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.ddrl_oracle as O  # noqa: E402

os.environ["DDRL_GN_STEP"] = "three-kernel"      # the orchestration checked here is the kernel-per-stage path (mocked below)

_dev = torch.device
torch.device = lambda *a, **k: _dev("cpu")
torch.cuda.is_available = lambda: True
torch.cuda.get_device_properties = lambda d: types.SimpleNamespace(multi_processor_count=148)
torch.cuda.synchronize = lambda *a: None
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.is_cuda = property(lambda self: True)

import ddrl_b200.kernels as K  # noqa: E402
import ddrl_b200.learner as L  # noqa: E402

D64 = lambda t: t.detach().double()       # noqa: E731


def graphnet_forward(theta, idx, st, adj, A, out=None):
    lg, v = O.graphnet_forward(D64(theta).reshape(-1), idx, D64(st), D64(adj), 2 * A)
    return lg.float(), v.float()


def dg_sample(logits, eps):
    a = O.dg_sample(D64(logits), D64(eps))
    return a.float(), O.dg_logp(D64(logits), a).float()


def gae(rewards, values, dones, v_boot, cpe, gamma, lam, adv=None, vtarg=None, moments=None, ws=None):
    P = rewards.shape[0]
    adv, vtarg = torch.empty_like(rewards), torch.empty_like(rewards)
    moments = torch.empty(P, 3, dtype=torch.float64)
    dn = np.repeat(dones.numpy(), cpe, axis=1)
    for p in range(P):
        a, v = O.gae_recurrence(rewards[p].numpy(), values[p].numpy(), dn, v_boot[p].numpy(), gamma, lam)
        adv[p], vtarg[p] = torch.from_numpy(a.astype(np.float32)), torch.from_numpy(v.astype(np.float32))
        a64 = a.astype(np.float32).astype(np.float64)
        moments[p] = torch.tensor([a64.size, a64.sum(), (a64 ** 2).sum()])
    return adv, vtarg, moments


def adv_standardize(adv, moments):
    P = adv.shape[0]
    flat = adv.view(P, -1)
    mean = moments[:, 1] / moments[:, 0]
    std = (moments[:, 2] / moments[:, 0] - mean ** 2).clamp_min(0).sqrt().clamp_min(1e-4)
    flat.copy_(((flat.double() - mean[:, None]) / std[:, None]).float())
    return adv


def ppo_loss_grad(logits, value, act, ol, olp, vfp, adv, vt, A, kl_coeff, hyper, LG, dlogits, dvalue, spart):
    with torch.enable_grad():
        lg, v = D64(logits[0]).requires_grad_(True), D64(value[0]).requires_grad_(True)
        logp = O.dg_logp(lg, D64(act))
        ratio = torch.exp(logp - D64(olp))
        kl, ent = O.dg_kl(D64(ol), lg), O.dg_entropy(lg)
        a = D64(adv)
        surr = torch.minimum(a * ratio, a * torch.clamp(ratio, 1 - hyper.clip_param, 1 + hyper.clip_param))
        vf1 = (v - D64(vt)) ** 2
        vf2 = (D64(vfp) + torch.clamp(v - D64(vfp), -hyper.vf_clip_param, hyper.vf_clip_param) - D64(vt)) ** 2
        vf = torch.maximum(vf1, vf2)
        total = ((-surr + float(kl_coeff[0]) * kl + hyper.vf_loss_coeff * vf - hyper.entropy_coeff * ent) * hyper.inv_global_mb).sum()
        gl, gv = torch.autograd.grad(total, (lg, v))
    dlogits.copy_(gl.float())
    dvalue.copy_(gv.float())
    R_, d_ = D64(vt), D64(vt) - v.detach()
    spart.zero_()
    spart[0, 0] = torch.stack([(-surr).sum(), kl.sum(), vf.sum(), ent.sum(), R_.sum(), (R_ * R_).sum(), d_.sum(), (d_ * d_).sum()]).detach()


def lib_backward(th, idx, st, adj, dlogits, dvalue, B, A, G, gpart):
    with torch.enable_grad():
        t = D64(th).requires_grad_(True)
        lg, v = O.graphnet_forward(t, idx, D64(st), D64(adj), 2 * A)
        (g,) = torch.autograd.grad((lg * D64(dlogits)).sum() + (v * D64(dvalue)).sum(), t)
    gpart.zero_()
    gpart[0, :g.numel()] = g.float()


def grad_reduce(gpart, spart, P, G, NP, grad, step_stats=None, step_ctr=None):
    grad.view(P, -1)[:, :NP] = gpart.reshape(P, G, -1)[:, :, :NP].sum(dim=1)
    if spart is not None and step_stats is not None:
        step_stats[int(step_ctr.item())] = spart.reshape(P, G, -1).sum(dim=1)


def clip_adam(theta, m, v, beta_pow, grad, lr, b1, b2, eps, clip, sync_ws, gnorm=None, step_ctr=None, img=None, img_D=0, img_A=0,
              tc_img=None):
    for p in range(theta.shape[0]):
        g = grad[p].double()
        norm = torch.sqrt((g * g).sum())
        g = g * (clip / torch.maximum(norm, torch.tensor(float(clip), dtype=torch.float64))) if clip > 0 else g
        if gnorm is not None:
            gnorm[p] = norm.float()
        b1p, b2p = float(beta_pow[p, 0]), float(beta_pow[p, 1])
        lr_t = lr * np.sqrt(1 - b2p) / (1 - b1p)
        m[p] = (b1 * m[p].double() + (1 - b1) * g).float()
        v[p] = (b2 * v[p].double() + (1 - b2) * g * g).float()
        theta[p] = (theta[p].double() - lr_t * m[p].double() / (v[p].double().sqrt() + eps)).float()
        beta_pow[p, 0] *= b1
        beta_pow[p, 1] *= b2
    if step_ctr is not None:
        step_ctr += 1


K.graphnet_forward, K.dg_sample, K.gae, K.adv_standardize = graphnet_forward, dg_sample, gae, adv_standardize
K.ppo_loss_grad, K.grad_reduce, K.clip_adam = ppo_loss_grad, grad_reduce, clip_adam
L._lib_backward = lib_backward

import tests.test_gpu_graphnet as T  # noqa: E402
import tests.test_gpu_graphnet_batch as Z  # noqa: E402

T.test_graphnet_ppo_iteration_runs_and_first_step_matches_oracle()
print("iteration ok")
Z.test_graphnet_learn_on_batch_is_the_sgd_phase_of_learn_on_rollout()
print("learn_on_batch ok")
