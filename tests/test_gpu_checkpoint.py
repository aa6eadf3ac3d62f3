"""GPU: a reference checkpoint (golden fixture of Results/**/checkpoint-1250) drives the learner — import, inference parity
with the oracle on the shipped weights + filter, continue training, export, re-import bit-exactly."""
import numpy as np
import pytest
import torch

from tests.test_checkpoint import _from_golden
from tests.util import scaled_err

pytestmark = pytest.mark.gpu


def _learner(ck, **kw):
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import FCNetLearner
    pc = next(iter(ck.policies.values()))
    return FCNetLearner(len(ck.policies), pc.obs_dim, pc.act_dim, PPOConfig(num_sgd_iter=1, sgd_minibatch_size=128), "cuda", **kw)


@pytest.mark.parametrize("arch", ["FullyDecentral", "TwoSides"])
def test_checkpoint_drives_the_learner(arch, tmp_path):
    import oracle.ddrl_oracle as O
    from ddrl_b200 import checkpoint as C
    ck = _from_golden(arch)
    L = _learner(ck)
    C.apply_to_learner(ck, L)
    pids = list(ck.policies)
    P, D, A = L.P, L.D, L.A
    # state landed
    for p, pid in enumerate(pids):
        pc = ck.policies[pid]
        assert np.array_equal(L.theta[p].cpu().numpy(), pc.theta)
        assert int(L.filt_n[p]) == pc.filter_n
        std = np.sqrt(pc.filter_S / (pc.filter_n - 1))
        np.testing.assert_allclose(L.norm[p, 0].cpu().numpy(), pc.filter_M, rtol=0, atol=0)
        np.testing.assert_allclose(L.norm[p, 1].cpu().numpy(), 1.0 / (std + 1e-8), rtol=1e-14)
        assert L.kl_coeff_host[p] == pytest.approx(0.45)
    # inference on the shipped weights + filter == oracle (float64), 1e-5 of the tensor's scale
    rng = np.random.default_rng(0)
    raw = np.stack([ck.policies[pid].filter_M + np.sqrt(ck.policies[pid].filter_S / (ck.policies[pid].filter_n - 1)) *
                    rng.standard_normal((64, D)) for pid in pids]).astype(np.float32)
    out = L.compute_actions(torch.from_numpy(raw).cuda())
    for p, pid in enumerate(pids):
        pc = ck.policies[pid]
        x = (raw[p].astype(np.float64) - pc.filter_M) / (np.sqrt(pc.filter_S / (pc.filter_n - 1)) + 1e-8)
        lg, v = O.fcnet_forward(torch.from_numpy(pc.theta).double(), torch.from_numpy(x.astype(np.float32)).double(), 2 * A)
        assert scaled_err(out["logits"][p].cpu().numpy(), lg.numpy()) < 1e-5
        assert scaled_err(out["value"][p].cpu().numpy(), v.numpy()) < 1e-5
    # continue training for one iteration, export, re-import into a fresh learner: bit-identical state
    T, Cc = 8, 32
    R = T * Cc
    rawb = torch.from_numpy(np.repeat(raw[:, None, :Cc], T, axis=1).copy()).cuda()
    rew = torch.from_numpy((0.3 + 0.5 * rng.standard_normal((P, T, Cc))).astype(np.float32)).cuda()
    dones = torch.zeros(T, Cc, dtype=torch.uint8, device="cuda")
    eps = torch.from_numpy(rng.standard_normal((P, T, Cc, A)).astype(np.float32)).cuda()
    perms = torch.from_numpy(np.stack([np.stack([rng.permutation(R // 128)]) for _ in range(P)]).astype(np.int32)).cuda()
    stats = L.learn_on_rollout(rawb, torch.from_numpy(raw[:, :Cc].copy()).cuda(), rew, dones, eps, perms)
    assert not np.array_equal(L.theta[0].cpu().numpy(), ck.policies[pids[0]].theta)
    ck2 = C.from_learner(L, pids, stats, {"num_steps_sampled": 123, "num_steps_trained": 123})
    path = str(tmp_path / "checkpoint-2")
    C.save_rllib_checkpoint(path, ck2)
    ck3 = C.load_rllib_checkpoint(path)
    L2 = _learner(ck3)
    C.apply_to_learner(ck3, L2)
    for name in ("theta", "m", "v", "beta_pow", "filt_n", "filt_M", "filt_S", "norm", "kl_coeff"):
        assert torch.equal(getattr(L, name), getattr(L2, name)), name
    assert ck3.counters["num_steps_trained"] == 123
    assert ck3.policies[pids[0]].learner_stats["cur_kl_coeff"] == pytest.approx(float(L.kl_coeff_host[0]))
