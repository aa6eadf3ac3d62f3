"""CPU: the DATAFLOW of the two-launch GraphNet SGD step (csrc/graphnet.cu `graphnet_train_fwd_kernel` /
`graphnet_train_acc_kernel`) restated in numpy float64 — per-row records {x_idx, mean of senders, dpre, y, dpx[node], dout},
then rank-1 weight-gradient accumulation with the hyper-network columns re-evaluated — against autograd through the oracle's
GraphNet.  It pins the formulas the kernels implement (which node receives which input gradient, the 1/cnt of the sender
mean, the tanh derivatives); it says nothing about the kernels' indexing or synchronisation (that is the opt-in GPU test)."""
import numpy as np
import pytest
import torch

import oracle.ddrl_oracle as O

H, F, E = 64, 19, 4


def _records_and_grads(theta, idx, state, adj, dout, n_out):
    p = {k: v.numpy() for k, v in O.unflatten(torch.from_numpy(theta), O.graphnet_shapes(n_out)).items()}
    We, be = p["state_enc/kernel"].reshape(E, F, H), p["state_enc/bias"].reshape(F, H)
    Wm, Wu, Wo = p["msg_transform/kernel"], p["node_update/kernel"], p["linear_out/kernel"]
    g = {k: np.zeros_like(v) for k, v in p.items()}
    gWe, gbe = g["state_enc/kernel"].reshape(E, F, H), g["state_enc/bias"].reshape(F, H)
    out = np.zeros((len(idx), n_out))
    for b in range(len(idx)):
        i = int(idx[b])
        snd = adj[b][:, i] != 0
        need = snd.copy()
        need[i] = True
        cnt = int(snd.sum())
        inv = 1.0 / cnt if cnt else 0.0
        # ---- launch 1: one row ----------------------------------------------------------------------------------------
        w = {n: np.tanh(be + np.einsum("q,qfh->fh", state[b, n, F:], We)) for n in range(4) if need[n]}
        x = {n: np.tanh(state[b, n, :F] @ w[n]) for n in w}
        xi = x[i]
        xm = sum((x[n] for n in range(4) if snd[n]), np.zeros(H)) * inv
        y = np.tanh(xi @ Wu + xm @ Wm)
        out[b] = y @ Wo + p["linear_out/bias"]
        d = dout[b]
        dpre = (Wo @ d) * (1 - y * y)
        dxi, dxm = Wu @ dpre, (Wm @ dpre) * inv
        dpx = {n: ((dxi if n == i else 0.0) + (dxm if snd[n] else 0.0)) * (1 - x[n] * x[n]) for n in w}
        # ---- launch 2: accumulate ----------------------------------------------------------------------------------------
        g["node_update/kernel"] += np.outer(xi, dpre)
        g["msg_transform/kernel"] += np.outer(xm, dpre)
        g["linear_out/kernel"] += np.outer(y, d)
        g["linear_out/bias"] += d
        for n in w:
            dpw = dpx[n][None, :] * state[b, n, :F][:, None] * (1 - w[n] * w[n])       # [F, H]
            gWe += state[b, n, F:][:, None, None] * dpw[None]
            gbe += dpw
    return out, np.concatenate([g[k].reshape(-1) for k, _ in O.graphnet_shapes(n_out)])


@pytest.mark.parametrize("adj_kind,n_out", [("ring", 4), ("random", 4), ("random", 1)])
def test_record_algorithm_equals_autograd(adj_kind, n_out):
    rng = np.random.default_rng(3 + n_out)
    B = 40
    state = rng.standard_normal((B, 4, 23))
    state[..., F:] = rng.uniform(-1, 1, (B, 4, 4))
    idx = rng.integers(0, 4, B)
    if adj_kind == "ring":
        adj = np.broadcast_to(O.ring_adjacency(torch.float64).numpy(), (B, 4, 4)).copy()
    else:      # weighted entries, self loops, receivers without senders
        adj = (rng.random((B, 4, 4)) < 0.4) * rng.uniform(0.5, 2.0, (B, 4, 4))
    gen = torch.Generator().manual_seed(1)
    theta = O.graphnet_init(n_out, gen, dtype=torch.float64)
    theta = theta + 0.05 * torch.randn(theta.shape, generator=gen, dtype=torch.float64)
    dout = rng.standard_normal((B, n_out))
    out, grad = _records_and_grads(theta.numpy(), idx, state, adj, dout, n_out)
    t = theta.clone().requires_grad_(True)
    ref_out = O.graphnet_forward_one(t, torch.from_numpy(idx), torch.from_numpy(state), torch.from_numpy(adj), n_out)
    (ref,) = torch.autograd.grad((ref_out * torch.from_numpy(dout)).sum(), t)
    assert np.abs(out - ref_out.detach().numpy()).max() < 1e-12
    assert np.abs(grad - ref.numpy()).max() < 1e-11 * max(1.0, np.abs(ref.numpy()).max())
