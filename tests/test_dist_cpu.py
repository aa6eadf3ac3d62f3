"""CPU, world_size 2, gloo: the host-side logic of the data-parallel learner (shard by env; gradient all-reduce with
1/global_minibatch scaling; rank-ordered filter merge; advantage-moment all-reduce) reproduces the single-process
result.  The per-rank compute is the oracle here (the CUDA kernels need a GPU; their N-GPU == 1-GPU test is
tests/test_gpu_multi.py, run with gpurun --gpus 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle.ddrl_oracle as O
from ddrl_b200.sharding import allreduce_sum_, gather_parts_rank_order, gather_stats_rank_order, local_minibatch, shard_envs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    g = torch.Generator().manual_seed(0)
    D, A, T, N = 19, 2, 8, 16
    th = O.fcnet_init(D, 2 * A, g, dtype=torch.float64)
    x = torch.randn(T, N, D, dtype=torch.float64, generator=g) * 3 + 1
    lg, v = O.fcnet_forward(th, x.reshape(-1, D), 2 * A)
    act = O.dg_sample(lg, torch.randn(T * N, A, dtype=torch.float64, generator=g)).detach()
    old = (lg + 0.05 * torch.randn(lg.shape, dtype=torch.float64, generator=g)).detach()
    adv = torch.randn(T, N, dtype=torch.float64, generator=g)
    vt = v.detach().reshape(T, N) + torch.randn(T, N, dtype=torch.float64, generator=g)
    return dict(D=D, A=A, T=T, N=N, th=th, x=x, act=act.reshape(T, N, A), old=old.reshape(T, N, 2 * A),
                old_logp=O.dg_logp(old, act).detach().reshape(T, N), vf=v.detach().reshape(T, N), adv=adv, vt=vt)


def _grad_sum(pr, env_slice, inv_global_rows):
    """sum over the local rows of d(row loss)/d(theta), scaled by 1/global rows (what a rank's kernel produces)."""
    cfg = O.PPOConfig()
    th = pr["th"].clone().requires_grad_(True)
    sl = lambda a: a[:, env_slice].reshape(-1, *a.shape[2:])
    lg, v = O.fcnet_forward(th, sl(pr["x"]), 2 * pr["A"])
    loss, _ = O.ppo_loss_from_outputs(lg, v, sl(pr["act"]), sl(pr["old"]), sl(pr["old_logp"]), sl(pr["vf"]),
                                      sl(pr["adv"]), sl(pr["vt"]), 0.2, cfg)
    n_local = lg.shape[0]
    (g,) = torch.autograd.grad(loss * n_local * inv_global_rows, th)
    return g


def _parts(x2d):
    """{count, mean, M2} per feature, the layout ddrl_filter_partial writes: [P=1, n=1, D, 3]."""
    rs = O.batch_stat(x2d.numpy())
    return torch.from_numpy(np.stack([np.full(x2d.shape[1], float(rs._n)), rs._M, rs._S], axis=-1))[None, None]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = _problem()
    lo, hi = shard_envs(pr["N"], world, rank)
    rows_global = pr["T"] * pr["N"]
    assert local_minibatch(rows_global, world) == pr["T"] * (hi - lo)
    g = allreduce_sum_(_grad_sum(pr, slice(lo, hi), 1.0 / rows_global), dist, world)
    mine = _parts(pr["x"][:, lo:hi].reshape(-1, pr["D"]))
    allp = gather_stats_rank_order(mine, dist, world)          # the product path: one merged triple per rank
    assert torch.equal(allp, gather_parts_rank_order(mine, dist, world))
    a = pr["adv"][:, lo:hi].reshape(-1)
    mom = allreduce_sum_(torch.stack([torch.tensor(float(a.numel()), dtype=torch.float64), a.sum(), (a * a).sum()]), dist, world)
    if rank == 0:
        out.put((g.numpy(), allp.numpy(), mom.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_host_logic_reproduces_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, allp, mom = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pr = _problem()
    # gradient: all-reduced sum of rank partials == gradient of the mean loss over the global minibatch
    g_ref = _grad_sum(pr, slice(0, pr["N"]), 1.0 / (pr["T"] * pr["N"])).numpy()
    np.testing.assert_allclose(g, g_ref, rtol=1e-10, atol=1e-14)
    # filter: partials arrive in rank order; merging them equals one pass over all rows
    assert allp.shape == (1, 2, pr["D"], 3)
    rs = O.RunningStat((pr["D"],))
    for i in range(allp.shape[1]):
        part = O.RunningStat((pr["D"],))
        part._n, part._M, part._S = int(allp[0, i, 0, 0]), allp[0, i, :, 1].copy(), allp[0, i, :, 2].copy()
        rs.update(part)
    full = O.batch_stat(pr["x"].reshape(-1, pr["D"]).numpy())
    assert rs.n == full.n == pr["T"] * pr["N"]
    np.testing.assert_allclose(rs._M, full._M, rtol=1e-12)
    np.testing.assert_allclose(rs._S, full._S, rtol=1e-11)
    first = O.batch_stat(pr["x"][:, :pr["N"] // 2].reshape(-1, pr["D"]).numpy())
    np.testing.assert_allclose(allp[0, 0, :, 1], first._M, rtol=1e-13)          # rank 0's partial comes first
    # advantage moments
    a = pr["adv"].reshape(-1)
    np.testing.assert_allclose(mom, [a.numel(), float(a.sum()), float((a * a).sum())], rtol=1e-12)


def test_shard_helpers():
    assert shard_envs(4096, 8, 3) == (1536, 2048)
    assert local_minibatch(4096, 8) == 512
    for bad in (lambda: shard_envs(10, 4, 0), lambda: local_minibatch(130, 4)):
        try:
            bad()
            raise AssertionError("expected ValueError")
        except ValueError:
            pass
