"""Diagnostic (not a test): per architecture and per variable, the gradient error of the tensor-core step (mode="tc", fp16
hi/lo split GEMMs), of the FP32-FMA step (mode="fp32") and of the oracle's float32 torch twin, all against the float64
oracle, on the published checkpoint weights.  Errors are max |g - g64| relative to the WHOLE gradient's scale of the policy
(max |g64|), the quantity the 1e-5 / 5e-5 statements in DESIGN.md §2 refer to.      python tests/grad_error_table.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tests.test_gpu_fcnet import _cuda_train_step, _make_batch, _oracle, _oracle_grads
from tests.util import load_ckpt

ARCHS = ["Centralized", "FullyDecentral", "Local", "SingleNeighbor", "SingleDiagonal", "SingleToFront", "TwoSides", "TwoDiags"]
NAMES = ["fc_1/kernel", "fc_1/bias", "fc_value_1/kernel", "fc_value_1/bias", "fc_2/kernel", "fc_2/bias", "fc_value_2/kernel",
         "fc_value_2/bias", "fc_out/kernel", "fc_out/bias", "value_out/kernel", "value_out/bias"]


def main():
    O = _oracle()
    cfg = O.PPOConfig(entropy_coeff=0.01)
    R, MB = 4096, 2048
    print(f"{'architecture':16s} {'variable':20s} {'tc':>10s} {'fp32':>10s} {'f32 twin':>10s}   (max |g - g64| / max |g64| over the policies)")
    worst = {"tc": 0.0, "fp32": 0.0, "twin": 0.0}
    for arch in ARCHS:
        b = _make_batch(arch, R, 3, "cuda")
        klc = [0.2] * b["P"]
        g_tc, _ = _cuda_train_step(b, MB, 1, 16, klc, cfg, tc=True)
        g_32, _ = _cuda_train_step(b, MB, 1, 16, klc, cfg, tc=False)
        ref, _ = _oracle_grads(b, slice(MB, 2 * MB), klc, cfg)
        twin, _ = _oracle_grads(b, slice(MB, 2 * MB), klc, cfg, torch.float32)
        z = load_ckpt(arch)
        shapes = z[[k for k in z.files if k.endswith("/shapes")][0]]
        rows = {}
        for p in range(b["P"]):
            scale = np.abs(ref[p]).max()
            o = 0
            for name, shp in zip(NAMES, shapes):
                n = int(shp[0] * max(1, shp[1]))
                e = lambda g: float(np.abs(np.asarray(g[p][o:o + n], dtype=np.float64) - ref[p][o:o + n]).max() / scale)
                r = rows.setdefault(name, [0.0, 0.0, 0.0])
                r[0], r[1], r[2] = max(r[0], e(g_tc)), max(r[1], e(g_32)), max(r[2], e(twin))
                o += n
        for name in NAMES:
            r = rows[name]
            print(f"{arch:16s} {name:20s} {r[0]:10.2e} {r[1]:10.2e} {r[2]:10.2e}")
            worst["tc"], worst["fp32"], worst["twin"] = max(worst["tc"], r[0]), max(worst["fp32"], r[1]), max(worst["twin"], r[2])
    print(f"\nworst over all architectures / variables: tc {worst['tc']:.2e}   fp32 {worst['fp32']:.2e}   float32 twin {worst['twin']:.2e}")


if __name__ == "__main__":
    main()
