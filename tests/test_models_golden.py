"""CPU: the oracle's model restatements against vectors produced by the REFERENCE'S OWN model code
(`/root/reference/models/*.py` imported and executed unmodified on a numpy stand-in for TensorFlow:
tests/golden/make_models_golden.py + tests/golden/tf_shim.py, whose header says what such vectors can and cannot pin).
float64 on both sides."""
import os

import numpy as np
import pytest
import torch

import oracle.ddrl_oracle as O
from tests.util import GOLDEN

G = np.load(os.path.join(GOLDEN, "models.npz"))
T = lambda k: torch.from_numpy(G[k])           # noqa: E731


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), np.abs(a - b).max()


def test_graph_ops():
    close(O.adj_norm(T("ops/adj")).numpy(), G["ops/adj_norm"])
    close(O.symm_norm(T("ops/adj")).numpy(), G["ops/symm_norm"])
    close(O.segment_softmax(T("ops/seg_data"), torch.from_numpy(G["ops/seg_ids"]), 5).numpy(), G["ops/segment_softmax"])


@pytest.mark.parametrize("adj", ["ring", "rand"])
def test_graph_layers(adj):
    x, a = T("layers/x"), T(f"layers/adj_{adj}")
    close(O.mpnn_layer(x, a, T("mpnn/W_msg"), T("mpnn/W_upd"), None, "tanh").numpy(), G[f"mpnn/y_{adj}"])
    close(O.mpnn2_layer(x, a, T("mpnn2/W_msg"), T("mpnn2/W_upd"), T("mpnn2/b"), "tanh").numpy(), G[f"mpnn2/y_{adj}"])
    close(O.gat1_layer(x, a, T("gat1/W_pre"), T("gat1/w_att"), T("gat1/b"), "tanh").numpy(), G[f"gat1/y_{adj}"])


def test_gcn_layer():
    close(O.gcn_layer(T("layers/x"), T("gcn/adj"), T("gcn/W"), T("gcn/b"), "tanh").numpy(), G["gcn/y"])


def test_graphnet_and_wrapper():
    idx, state, adj = torch.from_numpy(G["gn/idx"]), T("gn/state"), T("gn/adj")
    close(O.graphnet_forward_one(T("gn/theta"), idx, state, adj, 4).numpy(), G["gn/out"])
    logits, value = O.graphnet_forward(T("wrap/theta"), idx, state, adj, 4)
    close(logits.numpy(), G["wrap/logits"])
    close(value.numpy(), G["wrap/value"])
    assert G["wrap/theta"].size == O.n_params(O.graphnet_shapes(4)) + O.n_params(O.graphnet_shapes(1)) == 28869


@pytest.mark.parametrize("tag,vf_share,free_std", [("default", False, False), ("vfshare", True, False), ("freestd", False, True),
                                                   ("both", True, True)])
def test_fcnet_layouts(tag, vf_share, free_std):
    theta = T(f"fc/{tag}/theta")
    assert theta.numel() == O.n_params(O.fcnet_shapes(19, 4, (64, 64), vf_share, free_std))
    logits, value = O.fcnet_forward(theta, T("fc/x"), 4, vf_share_layers=vf_share, free_log_std=free_std)
    close(logits.numpy(), G[f"fc/{tag}/logits"])
    close(value.numpy(), G[f"fc/{tag}/value"])


def test_leg_coupling_layer_and_model():
    close(O.leg_coupling(T("cup/logits_in"), torch.from_numpy(G["cup/node_id"]), T("cup/coupling")).numpy(), G["cup/layer_out"])
    assert np.array_equal(G["cup/coupling"], np.asarray(O.COUPLING_INIT))
    logits, value = O.fcnet_forward(T("cupnet/theta"), T("cupnet/x"), 4)
    close(O.leg_coupling(logits, torch.from_numpy(G["cupnet/node_id"]), T("cup/coupling")).numpy(), G["cupnet/logits"])
    close(value.numpy(), G["cupnet/value"])
