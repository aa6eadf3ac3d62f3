"""Diagnostic (not a test): where a learner iteration of the bench workload spends its time — device-bound time of the
preparation phase and of the SGD phase (enqueued back to back, no host read in between) against the full iteration with its
host round trip (stats + KL update).      python tests/iter_breakdown.py   (on a B200)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl_b200 import kernels as K
from ddrl_b200.config import PPOConfig
from ddrl_b200.learner import FCNetLearner
import bench


def main():
    dev = torch.device("cuda", 0)
    W = bench.WORKLOAD
    P, D, A, T, E = W["P"], W["D"], W["A"], W["T"], W["epochs"]
    envs, nb = 4096, 32
    R = T * envs
    cfg = PPOConfig(num_sgd_iter=E, sgd_minibatch_size=R // nb)
    from ddrl_b200.modelv2 import fcnet_init_flat
    gen = torch.Generator().manual_seed(1)
    theta0 = torch.stack([fcnet_init_flat(D, 2 * A, gen) for _ in range(P)])
    L = FCNetLearner(P, D, A, cfg, dev, theta=theta0)
    s = bench.synth_rollout(P, T, envs, D, A, envs, nb, E, 3, device=dev)
    args = (s["raw"], s["boot"], s["rewards"], s["dones"], s["eps"], s["perms"], s["shuffle"])
    for _ in range(4):
        L.learn_on_rollout(*args)
    torch.cuda.synchronize()

    def timed(fn, n=10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    full = timed(lambda: L.learn_on_rollout(*args))
    b = L._alloc(T, envs)
    obs_flat, eps_flat = s["raw"].reshape(P, R, D), s["eps"].reshape(P, R, A)
    names = ("obs", "act", "logits", "logp", "value", "adv", "vtarg")
    src = {n_: b[n_ + "_s"] for n_ in names}
    prep = timed(lambda: L._prepare_cached(b, obs_flat, s["boot"], s["rewards"], s["dones"], eps_flat, s["shuffle"], 1, True, T, envs))
    MB, nbb, G = L._sgd_setup(R)
    hyper = L._hyper(MB)

    def sgd_only():
        L.step_ctr.zero_()
        L._sgd_step(b, MB, G, hyper, src, nsteps=E * nb)
    sgd = timed(sgd_only)

    def both():
        L._prepare_cached(b, obs_flat, s["boot"], s["rewards"], s["dones"], eps_flat, s["shuffle"], 1, True, T, envs)
        sgd_only()
    bb = timed(both)
    print(f"full learn_on_rollout      : {full[0]:.3f} ms device-event / {full[1]:.3f} ms wall per iteration")
    print(f"prepare phase only (graph) : {prep[0]:.3f} ms")
    print(f"SGD phase only             : {sgd[0]:.3f} ms")
    print(f"prepare + SGD, no host read: {bb[0]:.3f} ms   -> host round trip exposed per iteration: {full[0] - bb[0]:.3f} ms")


if __name__ == "__main__":
    main()
