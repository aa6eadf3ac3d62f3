"""GPU: `GraphNetLearner.learn_on_batch` on postprocessed sample-batch columns reproduces the SGD phase of
`learn_on_rollout` (SURVEY.md §8-f N4 for the shared graph policy)."""
import os

import numpy as np
import pytest
import torch

from tests.test_gpu_graphnet import _dev, _inputs, _theta

pytestmark = pytest.mark.gpu


def test_graphnet_learn_on_batch_is_the_sgd_phase_of_learn_on_rollout():
    from ddrl_b200 import kernels as K
    from ddrl_b200.config import PPOConfig
    from ddrl_b200.learner import GraphNetLearner
    A, T, N, E, NB = 2, 8, 16, 2, 4
    C, R = N * 4, 8 * 16 * 4
    cfg = PPOConfig(num_sgd_iter=E, sgd_minibatch_size=R // NB)
    idx, state, adj = _inputs(R, 31, "ring")
    idx = np.tile(np.arange(4, dtype=np.int32), R // 4)
    bidx, bstate, badj = _inputs(C, 32, "ring")
    rng = np.random.default_rng(33)
    rewards = (0.3 + 0.5 * rng.standard_normal((T, C))).astype(np.float32)
    dones = (rng.random((T, N)) < 0.05).astype(np.uint8)
    eps = rng.standard_normal((T, C, A)).astype(np.float32)
    perms = np.stack([rng.permutation(NB) for _ in range(E)]).astype(np.int32)
    shuffle = rng.permutation(R).astype(np.int32)
    th0 = _theta(A, 6, big=True).float()
    d = dict(idx=_dev(idx), st=_dev(state), adj=_dev(adj), bidx=_dev(bidx), bst=_dev(bstate), badj=_dev(badj),
             rew=_dev(rewards), dones=_dev(dones), eps=_dev(eps), perms=_dev(perms), shuffle=_dev(shuffle))

    La = GraphNetLearner(A, cfg, "cuda", theta=th0.reshape(1, -1))
    stats_a = La.learn_on_rollout(d["idx"].reshape(T, C), d["st"].reshape(T, C, 4, 23), d["adj"].reshape(T, C, 4, 4), d["bidx"],
                                  d["bst"], d["badj"], d["rew"], d["dones"], d["eps"], d["perms"], d["shuffle"])
    # the postprocessed columns a rollout worker would hand over, computed with the same kernels from the initial weights
    th = _dev(th0.numpy())
    logits, value = K.graphnet_forward(th, d["idx"], d["st"], d["adj"], A)
    act, logp = K.dg_sample(logits, d["eps"].reshape(R, A).contiguous())
    _, vboot = K.graphnet_forward(th, d["bidx"], d["bst"], d["badj"], A)
    adv, vtarg, _ = K.gae(d["rew"].reshape(1, T, C), value.reshape(1, T, C), d["dones"], vboot.reshape(1, C), 4, cfg.gamma,
                          cfg.lambda_)
    Lb = GraphNetLearner(A, cfg, "cuda", theta=th0.reshape(1, -1))
    stats_b = Lb.learn_on_batch(d["idx"], d["st"], d["adj"], act, logits, logp, value, adv.reshape(R), vtarg.reshape(R), d["perms"],
                                d["shuffle"], standardize=True)
    torch.cuda.synchronize()
    assert not torch.equal(La.theta.cpu(), th0.reshape(1, -1))
    assert torch.equal(La.theta, Lb.theta) and torch.equal(La.m, Lb.m) and torch.equal(La.v, Lb.v)
    assert np.array_equal(La.kl_coeff_host, Lb.kl_coeff_host) and stats_a == stats_b
    with pytest.raises(Exception, match="CUDA tensor"):
        Lb.learn_on_batch(d["idx"], d["st"], d["adj"], act, logits, logp, value[:-1], adv.reshape(R), vtarg.reshape(R), d["perms"])
