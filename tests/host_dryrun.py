"""Host-logic dry run (CPU, run in its OWN process by tests/test_rllib_policy.py — it monkeypatches torch globally).

The RLlib-facing surface (`learner.FCNetLearner.learn_on_batch`, `rllib_policy.PPOPolicyGroup`) is orchestration around
CUDA kernels.  Here every `ddrl_b200.kernels` entry it touches is replaced by the ORACLE's CPU restatement of the same
operation and the optimizer phase by the oracle's `sgd_loop`, "cuda" is mapped to "cpu", and the two GPU tests of
tests/test_gpu_rllib_policy.py are executed unchanged.  What this checks without a GPU: column stacking and validation,
the RandomState replay of the minibatch order, StandardizeFields before the ragged-tail truncation, the row order handed
to the optimizer phase, weights by TF variable name, error paths.  It checks NO kernel (that is the `-m gpu` suite) and is
test infrastructure only — the product has no CPU path."""
import os
import sys
import types

import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.ddrl_oracle as O

_dev = torch.device
torch.device = lambda *a, **k: _dev("cpu")
torch.cuda.is_available = lambda: True
torch.cuda.get_device_properties = lambda d: types.SimpleNamespace(multi_processor_count=148)
torch.cuda.synchronize = lambda *a: None
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.is_cuda = property(lambda self: True)

import ddrl_b200.kernels as K
import ddrl_b200.learner as L

K.fcnet_pack = lambda *a, **k: None
def tc_pack(theta, D, A, out=None):
    return out if out is not None else torch.zeros(theta.shape[0], 16, dtype=torch.uint8)
K.fcnet_tc_pack = tc_pack
K.tc_pingpong_eligible = lambda D, A: True
def fwd(theta, obs, A, norm=None, clip=0.0, eps=None, want_obs_out=False, out=None, img=None):
    P, R, D = obs.shape
    res = {"logits": torch.empty(P, R, 2 * A), "value": torch.empty(P, R)}
    if eps is not None:
        res["action"], res["logp"] = torch.empty(P, R, A), torch.empty(P, R)
    for p in range(P):
        x = obs[p]
        if norm is not None:
            x = ((x.double() - norm[p, 0]) * norm[p, 1]).float()
        lg, v = O.fcnet_forward(theta[p], x, 2 * A)
        res["logits"][p], res["value"][p] = lg, v
        if eps is not None:
            a = O.dg_sample(lg, eps[p]); res["action"][p] = a; res["logp"][p] = O.dg_logp(lg, a)
        if out and out.get("obs_out") is not None:
            out["obs_out"][p] = x
    if out:
        for k_, name in (("logits", "logits"), ("value", "value"), ("action", "action"), ("logp", "logp")):
            if out.get(k_) is not None and name in res:
                out[k_].copy_(res[name])
    return res
K.fcnet_forward = fwd
def fwd_tc(tc_img, obs, A, norm=None, clip=0.0, eps=None, out=None, status=None):
    return fwd(CUR["theta"], obs, A, norm=norm, eps=eps, out=out)
K.fcnet_forward_tc = fwd_tc
CUR = {}
def filt_update(x, n, M, S, norm, ws=None):
    norm[:, 0] = x.double().mean(1); norm[:, 1] = 1.0 / (x.double().std(1) + 1e-8)
K.filter_update = filt_update
def gae(rewards, values, dones, v_boot, cpe, gamma, lam, adv=None, vtarg=None, moments=None, ws=None):
    P = rewards.shape[0]
    adv = adv if adv is not None else torch.empty_like(rewards); vtarg = vtarg if vtarg is not None else torch.empty_like(rewards)
    moments = moments if moments is not None else torch.empty(P, 3, dtype=torch.float64)
    for p in range(P):
        a, v = O.gae_recurrence(rewards[p].numpy(), values[p].numpy(), dones.numpy(), v_boot[p].numpy(), gamma, lam)
        adv[p] = torch.from_numpy(np.asarray(a, dtype=np.float32)); vtarg[p] = torch.from_numpy(np.asarray(v, dtype=np.float32))
        moments[p] = torch.tensor([a.size, a.sum(), (a.astype(np.float64) ** 2).sum()])
    return adv, vtarg, moments
K.gae = gae
def adv_std(adv, moments):
    P = adv.shape[0]
    flat = adv.view(P, -1)
    mean = moments[:, 1] / moments[:, 0]; var = moments[:, 2] / moments[:, 0] - mean ** 2
    flat.copy_(((flat.double() - mean[:, None]) / torch.clamp(var.sqrt(), min=1e-4)[:, None]).float())
    return adv
K.adv_standardize = adv_std
def gather_rows(src, perm, dst):
    for p in range(src.shape[0]):
        dst[p] = src[p][perm[p].long()]
    return dst
K.gather_rows = gather_rows

def sgd_phase(self, b, src, src_key, perms, R, T, Cc):
    assert perms.shape[2] == max(1, R // min(self.cfg.sgd_minibatch_size, R)), perms.shape
    for k_, v in src.items():
        assert v.shape[1] == R, (k_, v.shape, R)
    fwd_ = lambda th, xx: O.fcnet_forward(th, xx, 2 * self.A)
    cfg_o = O.PPOConfig(num_sgd_iter=self.cfg.num_sgd_iter, sgd_minibatch_size=self.cfg.sgd_minibatch_size)
    stats = []
    for p in range(self.P):
        ob = {"obs": src["obs"][p], "actions": src["act"][p], "old_logits": src["logits"][p], "old_logp": src["logp"][p],
              "vf_preds": src["value"][p], "advantages": src["adv"][p], "value_targets": src["vtarg"][p]}
        if not hasattr(self, "_ost"):
            self._ost = [O.AdamState.zeros(self.NP, torch.float32, cfg_o) for _ in range(self.P)]
        th, st = O.sgd_loop(self.theta[p].clone(), self._ost[p], fwd_, ob, perms[p].numpy(), float(self.kl_coeff_host[p]), cfg_o)
        self.theta[p] = th
        st = dict(st); st["cur_kl_coeff"] = 0.2; st["cur_lr"] = 3e-4
        stats.append(st)
    self._update_kl(stats)
    return stats
L.FCNetLearner._sgd_phase = sgd_phase
_orig_prepare = L.FCNetLearner._prepare
def prepare(self, *a, **k):
    CUR["theta"] = self.theta
    return _orig_prepare(self, *a, **k)
L.FCNetLearner._prepare = prepare

import tests.test_gpu_rllib_policy as Tst
for mode in ("fp32", "tc"):
    Tst.test_learn_on_batch_is_the_sgd_phase_of_learn_on_rollout(mode)
    print("test1", mode, "ok")
    Tst.test_policy_group_serves_the_rllib_calls(mode)
    print("test2", mode, "ok")
