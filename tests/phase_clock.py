"""Diagnostic (not a test): in-kernel phase timing of the ping-pong tcgen05 step through ddrl_tc_set_debug_clock.
    python tests/phase_clock.py   (on a B200)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ddrl_b200 import _lib, kernels as K
from ddrl_b200.config import PPOConfig
from ddrl_b200.learner import FCNetLearner
import bench

NAMES = {0: "entry", 1: "setup done", 2: "inputs landed", 3: "x split+publish", 4: "F1(0) ready", 5: "tanh1(0)", 6: "publish",
         7: "F1(1) ready", 8: "tanh1(1)", 9: "publish", 10: "F2(0) ready", 11: "tanh2(0)", 12: "publish", 13: "F2(1) ready",
         14: "tanh2(1)", 15: "publish", 16: "heads ready", 17: "loss", 18: "publish", 19: "B1(0) ready", 20: "dz2(0)",
         21: "publish", 22: "B1(1) ready", 23: "dz2(1)", 24: "publish", 25: "B3(0) ready", 26: "dz1(0)", 27: "publish",
         28: "B3(1) ready", 29: "dz1(1)", 30: "publish", 31: "B5 done", 32: "write-out", 33: "stats+dealloc", 34: "tail end", 35: "image issued", 36: "step/perm read", 37: "prefetch fn setup",
         38: "tmem alloc", 39: "inputs issued", 40: "tail: barrier A", 41: "tail: slice reduce", 42: "tail: sq + barrier B",
         43: "tail: norm + Adam", 44: "tail: ticket", 45: "write-out entry barrier",
         46: "early TMEM->smem", 47: "early copy-out", 48: "late copy-out", 49: "early copy-out (loop)",
         50: "copy it0", 51: "copy it1", 52: "copy it2", 53: "copy it3", 54: "copy it4"}
ORDER = [0, 35, 36, 37, 38, 39] + list(range(1, 31)) + [46, 50, 51, 52, 53, 54, 49, 47, 31, 45, 32, 48, 33, 40, 41, 42, 43, 44, 34]


def main():
    dev = torch.device("cuda", 0)
    W = bench.WORKLOAD
    P, D, A, T, E = W["P"], W["D"], W["A"], W["T"], 2
    envs, nb = 4096, 32
    C = envs
    R = T * C
    cfg = PPOConfig(num_sgd_iter=E, sgd_minibatch_size=R // nb)
    import oracle.ddrl_oracle as O
    gen = torch.Generator().manual_seed(1)
    theta0 = torch.stack([O.fcnet_init(D, 2 * A, gen) for _ in range(P)])
    L = FCNetLearner(P, D, A, cfg, dev, theta=theta0, use_graph=False)
    s = bench.synth_rollout(P, T, C, D, A, envs, nb, E, 3, device=dev)
    clk = torch.zeros(64 + 8 * 160, dtype=torch.int64, device=dev)
    L.learn_on_rollout(s["raw"], s["boot"], s["rewards"], s["dones"], s["eps"], s["perms"], s["shuffle"])
    torch.cuda.synchronize()
    _lib.load().ddrl_tc_set_debug_clock(clk.data_ptr())
    acc = []
    for it in range(3):
        L.learn_on_rollout(s["raw"], s["boot"], s["rewards"], s["dones"], s["eps"], s["perms"], s["shuffle"])
        torch.cuda.synchronize()
        acc.append(clk.cpu().numpy().copy())      # the stamps of the LAST step of the iteration
    _lib.load().ddrl_tc_set_debug_clock(None)
    c = acc[-1]
    t0 = c[0]
    prev = t0
    for i in ORDER:
        if c[i] == 0:
            continue
        print(f"{i:2d} {NAMES[i]:18s} +{(c[i] - prev) / 1.965e3:7.2f} us   @{(c[i] - t0) / 1.965e3:7.2f} us")
        prev = c[i]


    # all CTAs, globaltimer (ns): 0 step start, 1 main loop done, 2 write-out + stats done, 3 barrier A passed, 4 slice
    # reduce done, 5 barrier B passed, 6 Adam done, 7 step end
    nc = L.P * L._sgd_setup(R)[2]
    print('CTAs', nc, 'cluster size', K.tc_last_cluster())
    g = c[64:64 + 8 * nc].reshape(nc, 8).astype(np.float64)
    base = g[:, 0].min()
    names = ["start", "main done", "write-out", "barrier A", "slice red", "barrier B", "adam", "end"]
    print("per-CTA globaltimer (us after the earliest step start): min / median / max over all CTAs")
    for k in range(8):
        col = (g[:, k] - base) / 1e3
        print(f"  {names[k]:10s} {col.min():7.2f} {np.median(col):7.2f} {col.max():7.2f}   argmax CTA {int(col.argmax())}")
    # mean durations between consecutive stamps over all CTAs and the captured iterations (globaltimer ticks are coarse —
    # 32 ns to 1 us depending on the part — so medians are quantised; means over 128 CTAs x 3 captures are not)
    allg = np.stack([a[64:64 + 8 * nc].reshape(nc, 8).astype(np.float64) for a in acc])
    dm = np.diff(allg, axis=2).mean(axis=(0, 1)) / 1e3
    print("mean phase durations (us): " + "  ".join(f"{names[k + 1]} {dm[k]:.3f}" for k in range(7)) + f"  | step {dm.sum():.3f}")
    d = (g[:, 1] - g[:, 0]) / 1e3
    print("main loop duration per CTA: min %.2f med %.2f max %.2f; slowest CTAs %s" % (d.min(), np.median(d), d.max(), np.argsort(-d)[:8].tolist()))
    d2 = (g[:, 2] - g[:, 1]) / 1e3
    print("write-out duration per CTA: min %.2f med %.2f max %.2f" % (d2.min(), np.median(d2), d2.max()))


if __name__ == "__main__":
    main()
