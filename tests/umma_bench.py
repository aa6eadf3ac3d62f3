"""Diagnostic (not a test): cycles per tcgen05.mma for the shapes the training kernels use.  python tests/umma_bench.py
Columns: acc = independent accumulators the issuing thread cycles through, warps = concurrently issuing warps (own accumulators);
cycles are per MMA and per issuing thread (aggregate rate = warps / that)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddrl_b200 import _lib

lib = _lib.load()
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
st = torch.zeros(1, dtype=torch.int32, device="cuda")
print("   M    N a_mn b_mn reps ks elect acc warps unif | issue cyc/mma | total cyc/mma")
CASES = [(128, 64, 0, 0), (128, 16, 0, 0), (128, 256, 0, 0), (64, 16, 1, 1), (64, 64, 1, 1), (128, 128, 1, 1)]
for (M, N, amn, bmn) in CASES:
    for reps, ks, elect, nacc, lw, unif in ((32, 8, 0, 1, 0, 0), (32, 8, 0, 4, 0, 0), (32, 8, 0, 1, 2, 0),
                                            (32, 8, 0, 1, 0, 1), (32, 8, 0, 2, 0, 1), (32, 8, 0, 4, 0, 1), (32, 8, 0, 1, 1, 1), (32, 8, 0, 1, 2, 1),
                                            (32, 8, 0, 2, 2, 1), (4, 4, 0, 1, 0, 1), (1, 4, 0, 1, 0, 1), (1, 4, 0, 3, 0, 1)):
        if nacc * N > (128 if lw else 512):
            continue
        code = ks | (elect << 8) | ((nacc - 1) << 9) | (lw << 12) | (unif << 14)
        for _ in range(2):
            _lib.check(lib.ddrl_umma_bench(M, N, amn, bmn, reps, code, cyc.data_ptr(), st.data_ptr(), None), "umma_bench")
            torch.cuda.synchronize()
        c = cyc.cpu().tolist()
        n = reps * ks * nacc      # MMAs per issuing thread (the mode bits are NOT part of the count)
        print(f"{M:4d} {N:4d} {amn:4d} {bmn:4d} {reps:4d} {ks:2d} {elect:5d} {nacc:3d} {1 << lw:5d} {unif:4d} | {c[0] / n:10.1f}    | {c[1] / n:10.1f}   status={int(st)}")
