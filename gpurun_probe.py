import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import tests.test_gpu_fcnet as T
import oracle.ddrl_oracle as O
cfg = O.PPOConfig(entropy_coeff=0.01)
names = [n for n, _ in O.fcnet_shapes(19, 4)]
for arch, G in (("Centralized", 5), ("TwoSides", 3), ("FullyDecentral", 37)):
    b = T._make_batch(arch, 1000, 3, "cuda")
    klc = [0.2 * 1.5 ** p for p in range(b["P"])]
    ref, _ = T._oracle_grads(b, slice(500, 1000), klc, cfg)
    for tc in (False, True):
        grad, _ = T._cuda_train_step(b, 500, 1, G, klc, cfg, tc=tc)
        p = 0
        shapes = O.fcnet_shapes(b["D"], 2 * b["A"])
        o = 0
        gs = np.abs(ref[p]).max()
        line = []
        for (n, shp) in shapes:
            k = int(np.prod(shp))
            e = np.abs(grad[p][o:o+k].astype(np.float64) - ref[p][o:o+k]).max()
            line.append(f"{n.split('/')[0][:10]}{'k' if len(shp)==2 else 'b'}:{e/gs:.1e}/{e/np.abs(ref[p][o:o+k]).max():.1e}")
            o += k
        print(arch, "TC" if tc else "FP32", "gscale %.3g" % gs, " ".join(line))
