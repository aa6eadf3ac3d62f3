"""CPU: a literal transcription of the INDEX ARITHMETIC of the two-launch GraphNet SGD step
(csrc/graphnet.cu `graphnet_train_fwd_kernel`, `graphnet_train_acc_kernel`) — flat shared-memory offsets (padded strides 65 /
17), lane- and thread-ownership (`lane`, `lane + 32`; `h = tid & 63`, `fs = tid >> 6`), record offsets, gradient write-out —
executed sequentially in numpy float64 and compared with autograd through the oracle.  The kernels had not run on a GPU when
this was written; this test checks what a CPU can check about them (indices and formulas, not synchronisation)."""
import numpy as np
import pytest
import torch

import oracle.ddrl_oracle as O

GH, GF, GE, GS, GN, GK, GMAXO = 64, 19, 4, 23, 4, 5, 16
GWS, GWF_WO = GH + 1, GMAXO + 1
XI, XM, DPRE, Y, DPX, DOUT, REC = 0, GH, 2 * GH, 3 * GH, 4 * GH, 8 * GH, 8 * GH + GMAXO


def offsets(n_out):
    o, p = {}, 0
    for k, n in (("We", GE * GF * GH), ("be", GF * GH), ("Wm", GH * GH), ("Wu", GH * GH), ("Wo", GH * n_out), ("bo", n_out)):
        o[k] = p
        p += n
    o["NP"] = p
    return o


def launch1_row(th, o, n_out, row, idx, adj_row, dout_fn):
    """One warp, one row: lanes vectorised as numpy axis 0 (lane = 0..31)."""
    lane = np.arange(32)
    sWe, sbe = th[o["We"]:o["We"] + GE * GF * GH], th[o["be"]:o["be"] + GF * GH]
    sWu, sWm, sWo = np.zeros(GH * GWS), np.zeros(GH * GWS), np.zeros(GH * GWF_WO)
    for i in range(GH * GH):
        r, c = i >> 6, i & 63
        sWu[r * GWS + c], sWm[r * GWS + c] = th[o["Wu"] + i], th[o["Wm"] + i]
    for i in range(GH * GMAXO):
        hh, q = divmod(i, GMAXO)
        sWo[hh * GWF_WO + q] = th[o["Wo"] + hh * n_out + q] if q < n_out else 0.0
    sbo = np.array([th[o["bo"] + q] if q < n_out else 0.0 for q in range(GMAXO)])
    sRow = np.zeros(128)
    sRow[:GN * GS] = row.reshape(-1)
    sRow[GN * GS:GN * GS + GN * GN] = adj_row.reshape(-1)
    snd = [sRow[GN * GS + n * GN + idx] != 0.0 for n in range(GN)]
    need = [snd[n] or n == idx for n in range(GN)]
    cnt = sum(snd)
    sX = np.zeros(GN * GH)
    a0, a1 = np.zeros((GN, 32)), np.zeros((GN, 32))
    for f in range(GF):
        j = f * GH + lane
        for n in range(GN):
            if need[n]:
                e = sRow[n * GS + GF:n * GS + GF + GE]
                p0 = sbe[j] + sum(e[q] * sWe[q * GF * GH + j] for q in range(GE))
                p1 = sbe[j + 32] + sum(e[q] * sWe[q * GF * GH + j + 32] for q in range(GE))
                sf = sRow[n * GS + f]
                a0[n] += sf * np.tanh(p0)
                a1[n] += sf * np.tanh(p1)
    for n in range(GN):
        if need[n]:
            sX[n * GH + lane], sX[n * GH + lane + 32] = np.tanh(a0[n]), np.tanh(a1[n])
    inv = 1.0 / cnt if cnt > 0 else 0.0
    m0 = sum((sX[n * GH + lane] for n in range(GN) if snd[n]), np.zeros(32))
    m1 = sum((sX[n * GH + lane + 32] for n in range(GN) if snd[n]), np.zeros(32))
    sRow[lane], sRow[lane + 32] = m0 * inv, m1 * inv
    xs = sX[idx * GH:(idx + 1) * GH].copy()
    y0 = sum(xs[hh] * sWu[hh * GWS + lane] + sRow[hh] * sWm[hh * GWS + lane] for hh in range(GH))
    y1 = sum(xs[hh] * sWu[hh * GWS + lane + 32] + sRow[hh] * sWm[hh * GWS + lane + 32] for hh in range(GH))
    y0, y1 = np.tanh(y0), np.tanh(y1)
    sO = np.array([(y0 * sWo[lane * GWF_WO + q] + y1 * sWo[(lane + 32) * GWF_WO + q]).sum() + sbo[q] for q in range(n_out)])
    sDo = dout_fn(sO)
    dy0 = sum(sDo[q] * sWo[lane * GWF_WO + q] for q in range(n_out))
    dy1 = sum(sDo[q] * sWo[(lane + 32) * GWF_WO + q] for q in range(n_out))
    dp0, dp1 = dy0 * (1 - y0 * y0), dy1 * (1 - y1 * y1)
    rr = np.full(REC, np.nan)
    rr[XI + lane], rr[XI + lane + 32] = xs[lane], xs[lane + 32]
    rr[XM + lane], rr[XM + lane + 32] = sRow[lane], sRow[lane + 32]
    rr[DPRE + lane], rr[DPRE + lane + 32] = dp0, dp1
    rr[Y + lane], rr[Y + lane + 32] = y0, y1
    rr[DOUT:DOUT + GMAXO] = [sDo[q] if q < n_out else 0.0 for q in range(GMAXO)]
    sRow[64 + lane], sRow[96 + lane] = dp0, dp1
    dxi0 = sum(sWu[lane * GWS + hh] * sRow[64 + hh] for hh in range(GH))
    dxi1 = sum(sWu[(lane + 32) * GWS + hh] * sRow[64 + hh] for hh in range(GH))
    dxm0 = sum(sWm[lane * GWS + hh] * sRow[64 + hh] for hh in range(GH)) * inv
    dxm1 = sum(sWm[(lane + 32) * GWS + hh] * sRow[64 + hh] for hh in range(GH)) * inv
    for n in range(GN):
        q0 = q1 = np.zeros(32)
        if need[n]:
            dx0 = (dxi0 if n == idx else 0.0) + (dxm0 if snd[n] else 0.0)
            dx1 = (dxi1 if n == idx else 0.0) + (dxm1 if snd[n] else 0.0)
            x0, x1 = sX[n * GH + lane], sX[n * GH + lane + 32]
            q0, q1 = dx0 * (1 - x0 * x0), dx1 * (1 - x1 * x1)
        rr[DPX + n * GH + lane], rr[DPX + n * GH + lane + 32] = q0, q1
    assert not np.isnan(rr).any()
    return sO, rr


def launch2_cta(th, o, n_out, rows, idxs, adjs, recs, bx, G):
    """One CTA of 256 threads: its partial gradient [NP] (threads vectorised as numpy axis 0)."""
    tid = np.arange(256)
    h, fs = tid & (GH - 1), tid >> 6
    f_k = [fs + 4 * k for k in range(GK)]
    ok = [f < GF for f in f_k]
    We = [[np.where(ok[k], th[o["We"] + q * GF * GH + np.minimum(f_k[k], GF - 1) * GH + h], 0.0) for q in range(GE)] for k in range(GK)]
    be = [np.where(ok[k], th[o["be"] + np.minimum(f_k[k], GF - 1) * GH + h], 0.0) for k in range(GK)]
    gWe = [[np.zeros(256) for _ in range(GE)] for _ in range(GK)]
    gbe = [np.zeros(256) for _ in range(GK)]
    gWu, gWm = [np.zeros(256) for _ in range(16)], [np.zeros(256) for _ in range(16)]
    gWo, gbo = [np.zeros(256) for _ in range(n_out)], np.zeros(256)
    for b in range(bx, len(rows), G):
        rr, idx = recs[b], int(idxs[b])
        dpre, y = rr[DPRE + h], rr[Y + h]
        for i in range(16):
            gWu[i] += rr[XI + 4 * i + fs] * dpre
            gWm[i] += rr[XM + 4 * i + fs] * dpre
        for q in range(n_out):
            gWo[q] += y * rr[DOUT + q]
        gbo += np.where(tid < n_out, rr[DOUT + np.minimum(tid, GMAXO - 1)], 0.0)
        for n in range(GN):
            if n == idx or adjs[b].reshape(-1)[n * GN + idx] != 0.0:
                dpx = rr[DPX + n * GH + h]
                st = rows[b].reshape(-1)[n * GS:(n + 1) * GS]
                e = st[GF:GF + GE]
                for k in range(GK):
                    pre = be[k] + sum(e[q] * We[k][q] for q in range(GE))
                    w = np.tanh(pre)
                    dpw = np.where(ok[k], dpx * st[np.minimum(f_k[k], GF - 1)] * (1 - w * w), 0.0)
                    for q in range(GE):
                        gWe[k][q] += e[q] * dpw
                    gbe[k] += dpw
    gp = np.zeros(o["NP"])
    for k in range(GK):
        m = ok[k]
        for q in range(GE):
            gp[o["We"] + q * GF * GH + f_k[k][m] * GH + h[m]] = gWe[k][q][m]
        gp[o["be"] + f_k[k][m] * GH + h[m]] = gbe[k][m]
    for i in range(16):
        gp[o["Wu"] + (4 * i + fs) * GH + h] = gWu[i]
        gp[o["Wm"] + (4 * i + fs) * GH + h] = gWm[i]
    z = fs == 0
    for q in range(n_out):
        gp[o["Wo"] + h[z] * n_out + q] = gWo[q][z]
    gp[o["bo"] + np.arange(n_out)] = gbo[:n_out]
    return gp


@pytest.mark.parametrize("n_out,adj_kind", [(4, "ring"), (1, "random"), (16, "random")])
def test_transcribed_kernels_equal_autograd(n_out, adj_kind):
    rng = np.random.default_rng(n_out)
    B, G = 7, 3
    state = rng.standard_normal((B, 4, 23))
    state[..., GF:] = rng.uniform(-1, 1, (B, 4, 4))
    idxs = rng.integers(0, 4, B)
    if adj_kind == "ring":
        adj = np.broadcast_to(O.ring_adjacency(torch.float64).numpy(), (B, 4, 4)).copy()
    else:
        adj = (rng.random((B, 4, 4)) < 0.45).astype(np.float64)
    gen = torch.Generator().manual_seed(2)
    theta = O.graphnet_init(n_out, gen, dtype=torch.float64)
    theta = theta + 0.05 * torch.randn(theta.shape, generator=gen, dtype=torch.float64)
    th = theta.numpy()
    o = offsets(n_out)
    assert o["NP"] == th.size
    dout = rng.standard_normal((B, n_out))
    outs, recs = [], []
    for b in range(B):
        out, rr = launch1_row(th, o, n_out, state[b], int(idxs[b]), adj[b], lambda s, b=b: dout[b])
        outs.append(out)
        recs.append(rr)
    grad = sum(launch2_cta(th, o, n_out, state, idxs, adj, recs, bx, G) for bx in range(G))
    t = theta.clone().requires_grad_(True)
    ref_out = O.graphnet_forward_one(t, torch.from_numpy(idxs), torch.from_numpy(state), torch.from_numpy(adj), n_out)
    (ref,) = torch.autograd.grad((ref_out * torch.from_numpy(dout)).sum(), t)
    assert np.abs(np.asarray(outs) - ref_out.detach().numpy()).max() < 1e-12
    assert np.abs(grad - ref.numpy()).max() < 1e-11 * max(1.0, np.abs(ref.numpy()).max())
