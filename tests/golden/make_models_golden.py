#!/usr/bin/env python
"""Golden vectors of the reference's MODEL code (rows a1, a10, a11, a12 and N1 of SURVEY.md §8) — its own Python executed
unmodified on a numpy stand-in for TensorFlow.

    python tests/golden/make_models_golden.py        (HERE only: needs /root/reference)

TensorFlow and Ray are not installed, so `tests/golden/tf_shim.py` is put into `sys.modules` under their names (numpy ops
with the documented semantics of the TF ops of the same name; read its header for what that does and does not pin), then
`/root/reference/models` is imported as a normal package and its classes are instantiated and called:
    models.graph_ops   adj_norm, symm_norm, segment_softmax
    models.gcn         GCN, MPNN, MPNN2, GAT1
    models.graph_net   GraphNet                      (hyper-network encoder -> MPNN -> gather -> linear_out)
    models.shared_graphnet_glorot_uniform_init       FullyConnectedNetwork_GNN_GlorotUniformInitializer (actor + critic)
    models.fcnet_glorot_uniform_init                 FullyConnectedNetwork_GlorotUniformInitializer, the default layout and
                                                     the vf_share_layers / free_log_std variants
    models.coupling_net_glorot_uniform_init          LegCoupling + the coupling ModelV2 class
Weights are seeded float64 arrays assigned to the layers BY NAME; inputs are seeded float64.  Writes tests/golden/models.npz
(inputs, weights in the oracle's flat order, outputs)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tf_shim  # noqa: E402

np.product = np.prod                       # removed in NumPy 2; the reference (2020) still calls it
tf_shim.install()
sys.path.insert(0, "/root/reference")
import models  # noqa: E402  (the reference's package: registers its four custom models with the stub catalog)
from models import graph_ops  # noqa: E402
from models.gcn import GCN, MPNN, MPNN2, GAT1  # noqa: E402
from models.graph_net import GraphNet  # noqa: E402
from models.coupling_net_glorot_uniform_init import LegCoupling  # noqa: E402
from ddrl_b200 import spaces  # noqa: E402

RNG = np.random.default_rng(20261018)
CFG = {"fcnet_activation": "tanh", "fcnet_hiddens": [64, 64], "no_final_linear": False, "vf_share_layers": False,
       "free_log_std": False}
OUT = {}


def rnd(*shape, scale=1.0):
    return RNG.standard_normal(shape) * scale


def ring():
    a = np.zeros((4, 4))
    for s, r in [(0, 1), (1, 2), (2, 3), (3, 0), (1, 0), (2, 1), (3, 2), (0, 3)]:
        a[s, r] = 1.0
    return a


def graph_inputs(B, F):
    x = rnd(B, 4, F)
    adj_ring = np.broadcast_to(ring(), (B, 4, 4)).copy()
    adj_rand = (RNG.random((B, 4, 4)) < 0.45).astype(np.float64)
    adj_rand[0] = 0.0                                   # a graph without edges: every receiver is an empty segment
    return x, adj_ring, adj_rand


def main():
    assert sorted(models.ModelCatalog.registered) == ["cup", "fc_glorot_uniform_init", "ffn", "gnn"]     # models/__init__.py:7-13

    # ---- graph_ops ---------------------------------------------------------------------------------------------------------
    adj = (RNG.random((6, 4, 4)) < 0.6).astype(np.float64) * RNG.uniform(0.5, 2.0, (6, 4, 4)) + np.eye(4)   # no empty rows
    OUT["ops/adj"], OUT["ops/adj_norm"], OUT["ops/symm_norm"] = adj, graph_ops.adj_norm(adj), graph_ops.symm_norm(adj)
    data, ids = rnd(11, 3), RNG.integers(0, 5, 11)
    OUT["ops/seg_data"], OUT["ops/seg_ids"] = data, ids
    OUT["ops/segment_softmax"] = graph_ops.segment_softmax(data, ids, 5)

    # ---- layers: GCN, MPNN, MPNN2, GAT1 ---------------------------------------------------------------------------------------
    B, F, U = 5, 19, 64
    x, adj_ring, adj_rand = graph_inputs(B, F)
    OUT["layers/x"], OUT["layers/adj_ring"], OUT["layers/adj_rand"] = x, adj_ring, adj_rand
    gcn = GCN(U, activation="tanh", use_bias=True)
    gcn(x, adj_ring + np.eye(4))                                   # build
    gcn.linear.kernel, gcn.bias = rnd(F, U, scale=0.3), rnd(U, scale=0.1)
    OUT["gcn/W"], OUT["gcn/b"], OUT["gcn/adj"] = gcn.linear.kernel, gcn.bias, adj_ring + np.eye(4)
    OUT["gcn/y"] = gcn(x, adj_ring + np.eye(4))
    mp = MPNN(U, activation="tanh", use_bias=False)
    mp(x, adj_ring)
    mp.msg_transform.kernel, mp.node_update.kernel = rnd(F, U, scale=0.3), rnd(F, U, scale=0.3)
    OUT["mpnn/W_msg"], OUT["mpnn/W_upd"] = mp.msg_transform.kernel, mp.node_update.kernel
    OUT["mpnn/y_ring"], OUT["mpnn/y_rand"] = mp(x, adj_ring), mp(x, adj_rand)
    mp2 = MPNN2(U, activation="tanh", use_bias=True)
    mp2(x, adj_ring)
    mp2.msg_transform.kernel, mp2.node_update.kernel, mp2.bias = rnd(2 * F, U, scale=0.3), rnd(F + U, U, scale=0.3), rnd(U, scale=0.1)
    OUT["mpnn2/W_msg"], OUT["mpnn2/W_upd"], OUT["mpnn2/b"] = mp2.msg_transform.kernel, mp2.node_update.kernel, mp2.bias
    OUT["mpnn2/y_ring"], OUT["mpnn2/y_rand"] = mp2(x, adj_ring), mp2(x, adj_rand)
    gat = GAT1(U, activation="tanh", use_bias=True)
    gat(x, adj_ring)
    gat.pre_att_linear.kernel, gat.att_linear.kernel, gat.bias = rnd(F, U, scale=0.3), rnd(2 * U, 1, scale=0.3), rnd(U, scale=0.1)
    OUT["gat1/W_pre"], OUT["gat1/w_att"], OUT["gat1/b"] = gat.pre_att_linear.kernel, gat.att_linear.kernel, gat.bias
    OUT["gat1/y_ring"], OUT["gat1/y_rand"] = gat(x, adj_ring), gat(x, adj_rand)

    # ---- GraphNet and the actor / critic wrapper ----------------------------------------------------------------------------------
    B = 9
    state = rnd(B, 4, 23)
    state[..., 19:] = RNG.uniform(-1, 1, (B, 4, 4))
    idx = RNG.integers(0, 4, (B, 1))
    adjg = np.broadcast_to(ring(), (B, 4, 4)).copy()
    adjg[1] = (RNG.random((4, 4)) < 0.5).astype(np.float64)
    OUT["gn/state"], OUT["gn/idx"], OUT["gn/adj"] = state, idx, adjg

    def load_graphnet(net, n_out):
        net(np.zeros((1, 1), np.int32), np.zeros((1, 4, 23)), np.zeros((1, 4, 4)))                       # build
        w = [rnd(4, 19 * 64, scale=0.4), rnd(19 * 64, scale=0.2), rnd(64, 64, scale=0.2), rnd(64, 64, scale=0.2),
             rnd(64, n_out, scale=0.2), rnd(n_out, scale=0.1)]
        net.enc.kernel, net.enc.bias, net.gnn.msg_transform.kernel, net.gnn.node_update.kernel, net.out.kernel, net.out.bias = w
        return np.concatenate([a.reshape(-1) for a in w])      # oracle order: state_enc(k,b), msg_transform, node_update, linear_out(k,b)

    gn = GraphNet(4, CFG)
    OUT["gn/theta"] = load_graphnet(gn, 4)
    OUT["gn/out"] = gn(idx, state, adjg)
    graph_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (4, 23), np.float64),
                                spaces.MultiDiscrete(np.ones([4, 4]) * 2)])
    wrap = models.FullyConnectedNetwork_GNN_GlorotUniformInitializer(graph_space, spaces.Box(-1, 1, (2,)), 4, CFG, "leg_policy")
    OUT["wrap/theta"] = np.concatenate([load_graphnet(wrap.actor, 4), load_graphnet(wrap.critic, 1)])
    logits, st = wrap.forward({"obs": (idx, state, adjg)}, [], None)
    assert st == [] and len(wrap.variables()) == 12
    OUT["wrap/logits"], OUT["wrap/value"] = logits, wrap.value_function()

    # the `_GlobalCritic` wrapper cannot be constructed as shipped (shared_graphnet_glorot_uniform_init.py:69: super() names a
    # class it does not derive from) — the reason DESIGN.md lists it as "no behaviour to be identical to"
    from models.shared_graphnet_glorot_uniform_init import FullyConnectedNetwork_GNN_GlorotUniformInitializer_GlobalCritic as GC
    try:
        GC(graph_space, spaces.Box(-1, 1, (2,)), 4, CFG, "x")
        raise AssertionError("the reference's GlobalCritic wrapper became constructible: mirror it")
    except TypeError as exc:
        assert "super(type, obj)" in str(exc)

    # the Decentral_Graph scope hands GraphNet node matrices [4, 19]: the encoder splits off the last 4 columns and multiplies
    # the remaining 15 with a [19, 64] matrix (graph_net.py:33-37) — the reference model cannot even be built on that space
    narrow = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (4, 19), np.float64),
                           spaces.MultiDiscrete(np.ones([4, 4]) * 2)])
    try:
        models.FullyConnectedNetwork_GNN_GlorotUniformInitializer(narrow, spaces.Box(-1, 1, (2,)), 4, CFG, "x")
        raise AssertionError("GraphNet accepted [4, 19] node matrices")
    except ValueError as exc:
        assert "19" in str(exc) and "15" in str(exc)

    # ---- FCNet: default layout and the two optional ones ------------------------------------------------------------------------------
    D, A, B = 19, 2, 13
    xf = rnd(B, D)
    OUT["fc/x"] = xf
    for tag, vf_share, free_std in (("default", False, False), ("vfshare", True, False), ("freestd", False, True), ("both", True, True)):
        cfg = dict(CFG, vf_share_layers=vf_share, free_log_std=free_std)
        m = models.FullyConnectedNetwork_GlorotUniformInitializer(spaces.Box(-np.inf, np.inf, (D,), np.float64),
                                                                   spaces.Box(-1, 1, (A,)), 2 * A, cfg, "p")
        n_out = A if free_std else 2 * A
        names = ["fc_1"] + ([] if vf_share else ["fc_value_1"]) + ["fc_2"] + ([] if vf_share else ["fc_value_2"]) + ["fc_out", "value_out"]
        assert sorted(l.name for l in m.base_model.layers if isinstance(l, tf_shim.Dense)) == sorted(names)
        flat = []
        if free_std:
            m.log_std_var[:] = rnd(A, scale=0.3)
            flat.append(m.log_std_var.astype(np.float64))
        for nme in names:        # the oracle's order: (log_std,) fc_1, [fc_value_1,] fc_2, [fc_value_2,] fc_out, value_out
            lay = m.base_model.get_layer(nme)
            fan_in = D if nme.endswith("_1") else 64
            units = {"fc_out": n_out, "value_out": 1}.get(nme, 64)
            assert lay.kernel.shape == (fan_in, units)
            lay.kernel, lay.bias = rnd(fan_in, units, scale=0.3), rnd(units, scale=0.1)
            flat += [lay.kernel.reshape(-1), lay.bias]
        out, st = m.forward({"obs_flat": xf}, [], None)
        OUT[f"fc/{tag}/theta"], OUT[f"fc/{tag}/logits"], OUT[f"fc/{tag}/value"] = np.concatenate(flat), out, m.value_function()

    # ---- LegCoupling and the coupling ModelV2 class -----------------------------------------------------------------------------------------
    lc = LegCoupling()
    lg = rnd(7, 4)
    nid = RNG.integers(0, 4, (7, 1))
    OUT["cup/logits_in"], OUT["cup/node_id"], OUT["cup/layer_out"] = lg, nid, lc(lg, nid)
    OUT["cup/coupling"] = lc.coupling.astype(np.float64)
    cup_space = spaces.Tuple([spaces.MultiDiscrete([4]), spaces.Box(-np.inf, np.inf, (D,), np.float64)])
    cup_space.original_space = cup_space
    cm = models.FullyConnectedNetwork_Coupling_GlorotUniformInitializer(cup_space, spaces.Box(-1, 1, (A,)), 2 * A, CFG, "policy_legs")
    flat = []
    for nme in ["fc_1", "fc_value_1", "fc_2", "fc_value_2", "fc_out", "value_out"]:
        lay = cm.base_model.get_layer(nme)
        lay.kernel, lay.bias = rnd(*lay.kernel.shape, scale=0.3), rnd(lay.kernel.shape[1], scale=0.1)
        flat += [lay.kernel.reshape(-1), lay.bias]
    xc, nidc = rnd(B, D), RNG.integers(0, 4, (B, 1))
    out, _ = cm.forward({"obs": (nidc, xc)}, [], None)
    OUT["cupnet/theta"], OUT["cupnet/x"], OUT["cupnet/node_id"] = np.concatenate(flat), xc, nidc
    OUT["cupnet/logits"], OUT["cupnet/value"] = out, cm.value_function()

    np.savez_compressed(os.path.join(HERE, "models.npz"), **{k: np.asarray(v) for k, v in OUT.items()})
    print("wrote models.npz:", len(OUT), "arrays")


if __name__ == "__main__":
    main()
