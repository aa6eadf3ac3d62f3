"""Diagnostic (not a test): cycles per tcgen05.mma for the shapes the training kernels use.  python tests/umma_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl_b200 import _lib

lib = _lib.load()
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.zeros(1, dtype=torch.int32, device="cuda")
print("   M    N a_mn b_mn  reps ks | issue cyc/mma | total cyc/mma")
for (M, N, amn, bmn) in [(128, 64, 0, 0), (128, 16, 0, 0), (128, 256, 0, 0), (64, 16, 1, 1), (64, 64, 1, 1), (128, 128, 1, 1)]:
    for reps, ks in ((32, 8), (32, 8 + 256)):
        for _ in range(2):
            _lib.check(lib.ddrl_umma_bench(M, N, amn, bmn, reps, ks, cyc.data_ptr(), st.data_ptr(), None), "umma_bench")
            torch.cuda.synchronize()
        c = cyc.cpu().tolist()
        n = reps * ks
        print(f"{M:4d} {N:4d} {amn:4d} {bmn:4d} {reps:5d} {ks:2d} | {c[0] / n:10.1f}    | {c[1] / n:10.1f}   status={int(st)}")
