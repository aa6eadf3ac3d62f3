"""RLlib ModelV2 custom models of the reference, backed by the sm_100a kernels.

Same class names, constructor signature, ``forward(input_dict, state, seq_lens) -> (model_out, state)``,
stateful ``value_function()``, ``register_variables`` / ``variables`` / ``trainable_variables`` and registry names as
``models/`` in the reference (models/__init__.py:7-13):

    "ffn", "fc_glorot_uniform_init" -> FullyConnectedNetwork_GlorotUniformInitializer   (models/fcnet_glorot_uniform_init.py)
    "gnn"                           -> FullyConnectedNetwork_GNN_GlorotUniformInitializer (models/shared_graphnet_glorot_uniform_init.py)
    "cup"                           -> FullyConnectedNetwork_Coupling_GlorotUniformInitializer (models/coupling_net_glorot_uniform_init.py)

Parameters live in ONE flat float32 CUDA tensor per model in the reference's variable order, so checkpoints
(Results/**/checkpoint-1250) import with a copy.  Forward and backward run in libddrl_b200.so through a
``torch.autograd.Function`` (gradients w.r.t. the parameters only — observations are data).  With ray installed the
classes additionally subclass ``TorchModelV2``."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import kernels as K
from ._lib import DDRLError
from .catalog import ModelCatalog

try:  # pragma: no cover - ray is absent in the build container
    from ray.rllib.models.torch.torch_modelv2 import TorchModelV2 as _Base  # type: ignore
    _HAVE_RAY = True
except Exception:
    _HAVE_RAY = False

    class _Base:  # minimal ModelV2 surface the reference's classes use
        def __init__(self, obs_space, action_space, num_outputs, model_config, name):
            self.obs_space, self.action_space = obs_space, action_space
            self.num_outputs, self.model_config, self.name = num_outputs, model_config, name
            self.var_list: List[torch.Tensor] = []

        def get_initial_state(self):
            return []

        def __call__(self, input_dict, state=None, seq_lens=None):
            d = dict(input_dict)
            if "obs_flat" not in d and torch.is_tensor(d.get("obs")):
                d["obs_flat"] = d["obs"].reshape(d["obs"].shape[0], -1)
            out, state = self.forward(d, state or [], seq_lens)
            self._last_output = out
            return out, state

        def from_batch(self, train_batch, is_training=True):
            return self.__call__({"obs": train_batch["obs"], "is_training": is_training})

        def last_output(self):
            return self._last_output


def glorot_uniform_scaled_(w: torch.Tensor, scale: float, gen: Optional[torch.Generator]) -> torch.Tensor:
    """GlorotUniformScaled (models/glorot_uniform_scaled_initializer.py:14-19): VarianceScaling(scale, 'fan_avg',
    'uniform') -> U(-L, L), L = sqrt(6*scale/(fan_in+fan_out)).  TF's Philox stream cannot be bit-matched; parity of
    the initialiser is distributional (tests check bounds and variance L^2/3)."""
    fan_in, fan_out = w.shape
    lim = math.sqrt(6.0 * scale / (fan_in + fan_out))
    w.copy_((torch.rand(w.shape, generator=gen, dtype=torch.float64) * 2.0 - 1.0) * lim)
    return w


class _FlatParams:
    """Named views into one flat parameter vector (reference variable order)."""

    def __init__(self, shapes: Sequence):
        self.shapes = list(shapes)
        self.numel = int(sum(int(np.prod(s)) for _, s in self.shapes))

    def init_host(self, gen, small: Sequence[str]) -> torch.Tensor:
        flat = torch.zeros(self.numel, dtype=torch.float32)
        o = 0
        for name, shp in self.shapes:
            n = int(np.prod(shp))
            if len(shp) == 2:
                scale = 0.01 if any(name.startswith(s) for s in small) else 1.0
                glorot_uniform_scaled_(flat[o:o + n].view(shp), scale, gen)
            o += n
        return flat

    def views(self, flat: torch.Tensor, prefix: str = "") -> "OrderedDict[str, torch.Tensor]":
        out, o = OrderedDict(), 0
        for name, shp in self.shapes:
            n = int(np.prod(shp))
            out[prefix + name] = flat[o:o + n].view(shp)
            o += n
        return out


def fcnet_shapes(D: int, num_outputs: int):
    H = K.HIDDEN
    return [("fc_1/kernel", (D, H)), ("fc_1/bias", (H,)), ("fc_value_1/kernel", (D, H)), ("fc_value_1/bias", (H,)),
            ("fc_2/kernel", (H, H)), ("fc_2/bias", (H,)), ("fc_value_2/kernel", (H, H)), ("fc_value_2/bias", (H,)),
            ("fc_out/kernel", (H, num_outputs)), ("fc_out/bias", (num_outputs,)),
            ("value_out/kernel", (H, 1)), ("value_out/bias", (1,))]


def fcnet_variant_shapes(D: int, num_outputs: int, vf_share_layers: bool = False, free_log_std: bool = False):
    """Variables of the FCNet in the order the reference creates them for its two optional layouts
    (models/fcnet_glorot_uniform_init.py): `free_log_std` registers a state-independent `log_std[A]` FIRST and halves the
    width of fc_out (:30-36, 85-93); `vf_share_layers` drops the fc_value_* layers — value_out then reads the policy
    branch's last hidden layer (:95-113)."""
    H = K.HIDDEN
    n_out = num_outputs // 2 if free_log_std else num_outputs
    out = [("log_std", (n_out,))] if free_log_std else []
    for i, fan_in in ((1, D), (2, H)):
        out += [(f"fc_{i}/kernel", (fan_in, H)), (f"fc_{i}/bias", (H,))]
        if not vf_share_layers:
            out += [(f"fc_value_{i}/kernel", (fan_in, H)), (f"fc_value_{i}/bias", (H,))]
    out += [("fc_out/kernel", (H, n_out)), ("fc_out/bias", (n_out,)), ("value_out/kernel", (H, 1)), ("value_out/bias", (1,))]
    return out


def fcnet_layout_map(D: int, num_outputs: int, vf_share_layers: bool = False, free_log_std: bool = False) -> torch.Tensor:
    """int64 [NP_kernel]: for every element of the kernels' parameter layout (`fcnet_shapes`, two separate 64-64 branches,
    fc_out of width 2A) the index of the model variable that supplies it, or n_model for a constant zero.

    The kernels always evaluate two branches and a state-dependent log-std; the optional layouts are the same arithmetic
    on tied / constant weights:
      vf_share_layers: fc_value_i := fc_i, so the "value branch" recomputes the policy branch's hidden layers and
                       value_out reads h2 — exactly the shared-layer network; gradients of the two copies add up;
      free_log_std:    fc_out/kernel = [W | 0], fc_out/bias = [b | log_std]: logits = [h2 W + b, log_std] for every row.
    `theta_kernel = cat(theta_model, [0])[map]`, so autograd's index-select backward performs the gradient tying."""
    model = fcnet_variant_shapes(D, num_outputs, vf_share_layers, free_log_std)
    off, o = {}, 0
    for name, shp in model:
        off[name] = o
        o += int(np.prod(shp))
    n_model = o
    H, n_out = K.HIDDEN, (num_outputs // 2 if free_log_std else num_outputs)
    parts = []
    for name, shp in fcnet_shapes(D, num_outputs):
        n = int(np.prod(shp))
        src = name.replace("fc_value_", "fc_") if vf_share_layers else name
        if free_log_std and name == "fc_out/kernel":
            idx = torch.full((H, num_outputs), n_model, dtype=torch.int64)
            idx[:, :n_out] = off[src] + torch.arange(H * n_out).reshape(H, n_out)
            parts.append(idx.reshape(-1))
        elif free_log_std and name == "fc_out/bias":
            parts.append(torch.cat([off[src] + torch.arange(n_out), off["log_std"] + torch.arange(n_out)]))
        else:
            parts.append(off[src] + torch.arange(n))
    return torch.cat(parts)


def fcnet_init_flat(D: int, num_outputs: int, gen: Optional[torch.Generator] = None) -> torch.Tensor:
    """Flat float32 parameter vector of one FCNet policy initialised like the reference (GlorotUniformScaled: scale 1.0
    for the hidden layers, 0.01 for fc_out / value_out, zero biases; models/fcnet_glorot_uniform_init.py:48-113)."""
    return _FlatParams(fcnet_shapes(D, num_outputs)).init_host(gen, small=("fc_out", "value_out"))


def graphnet_shapes(num_outputs: int):
    H, F, E = K.HIDDEN, K.GN_FEATS, K.GN_ENC_IN
    one = lambda O, pre: [(pre + "state_enc/kernel", (E, F * H)), (pre + "state_enc/bias", (F * H,)),
                          (pre + "msg_transform/kernel", (H, H)), (pre + "node_update/kernel", (H, H)),
                          (pre + "linear_out/kernel", (H, O)), (pre + "linear_out/bias", (O,))]
    return one(num_outputs, "actor/") + one(1, "critic/")


def _check_fcnet_config(model_config: dict, who: str):
    """The kernels implement hiddens [64, 64], tanh and a final linear layer (every published run,
    Results/**/params.json); `vf_share_layers` and `free_log_std` are served through `fcnet_layout_map`.  Anything else
    fails loudly."""
    hid = list(model_config.get("fcnet_hiddens", [64, 64]))
    act = model_config.get("fcnet_activation", "tanh")
    bad = []
    if hid != [64, 64]:
        bad.append(f"fcnet_hiddens={hid} (supported: [64, 64])")
    if act != "tanh":
        bad.append(f"fcnet_activation={act!r} (supported: 'tanh')")
    if model_config.get("no_final_linear"):
        bad.append("no_final_linear=True")
    if bad:
        raise DDRLError(f"{who}: unsupported model_config for the sm_100a kernels: " + "; ".join(bad))


class _FCNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, obs, A):
        res = K.fcnet_forward(theta.detach().reshape(1, -1), obs.reshape(1, *obs.shape), A)
        ctx.save_for_backward(theta, obs)
        ctx.A = A
        return res["logits"][0], res["value"][0]

    @staticmethod
    def backward(ctx, dlogits, dvalue):
        theta, obs = ctx.saved_tensors
        B = obs.shape[0]
        dl = (dlogits if dlogits is not None else torch.zeros(B, 2 * ctx.A, device=obs.device)).contiguous().float()
        dv = (dvalue if dvalue is not None else torch.zeros(B, device=obs.device)).contiguous().float()
        g = K.fcnet_backward(theta.detach().reshape(1, -1), obs.reshape(1, *obs.shape), dl.reshape(1, B, -1),
                             dv.reshape(1, B), ctx.A)
        return g.reshape(theta.shape), None, None


class _GraphNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, node_idx, state, adj, A):
        lg, v = K.graphnet_forward(theta.detach(), node_idx, state, adj, A)
        ctx.save_for_backward(theta, node_idx, state, adj)
        ctx.A = A
        return lg, v

    @staticmethod
    def backward(ctx, dlogits, dvalue):
        theta, node_idx, state, adj = ctx.saved_tensors
        B = state.shape[0]
        dl = (dlogits if dlogits is not None else torch.zeros(B, 2 * ctx.A, device=state.device)).contiguous().float()
        dv = (dvalue if dvalue is not None else torch.zeros(B, device=state.device)).contiguous().float()
        g = K.graphnet_backward(theta.detach(), node_idx, state, adj, dl, dv, ctx.A)
        return g, None, None, None, None


class _DDRLModel(_Base, torch.nn.Module if _HAVE_RAY else object):
    def _setup(self, obs_space, action_space, num_outputs, model_config, name):
        if _HAVE_RAY:  # pragma: no cover
            torch.nn.Module.__init__(self)
        _Base.__init__(self, obs_space, action_space, num_outputs, model_config, name)
        if not torch.cuda.is_available():
            raise DDRLError(f"{type(self).__name__}: a CUDA device is required (no CPU fallback)")
        self._value_out = None
        self._registered: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    # reference API (TFModelV2.register_variables / variables / trainable_variables)
    def register_variables(self, variables):
        items = variables.items() if isinstance(variables, dict) else [(f"var_{len(self._registered) + i}", v)
                                                                       for i, v in enumerate(variables)]
        for k, v in items:
            self._registered[k] = v

    def variables(self, as_dict: bool = False):
        return OrderedDict(self._registered) if as_dict else list(self._registered.values())

    def trainable_variables(self, as_dict: bool = False):
        return self.variables(as_dict)

    def value_function(self):
        if self._value_out is None:
            raise DDRLError("value_function() called before forward()")
        return self._value_out.reshape(-1)

    # flat-parameter access for the fused learner / checkpoint import
    def flat_parameters(self) -> torch.Tensor:
        return self.theta

    def load_flat(self, flat):
        t = torch.as_tensor(flat, dtype=torch.float32)
        if t.numel() != self.theta.numel():
            raise DDRLError(f"load_flat: {t.numel()} values for a model with {self.theta.numel()} parameters")
        with torch.no_grad():
            self.theta.copy_(t.reshape(self.theta.shape).to(self.theta.device))


class FullyConnectedNetwork_GlorotUniformInitializer(_DDRLModel):
    """tanh MLP policy head + parallel value MLP, Glorot-uniform init with 0.01-scaled output layers
    (models/fcnet_glorot_uniform_init.py:17-125)."""

    def __init__(self, obs_space, action_space, num_outputs, model_config, name):
        self._setup(obs_space, action_space, num_outputs, model_config, name)
        _check_fcnet_config(model_config, type(self).__name__)
        self.D = int(np.prod(obs_space.shape))
        if num_outputs % 2:
            raise DDRLError("num_outputs must be 2 * action_dim (DiagGaussian)")
        self.A = num_outputs // 2
        self.vf_share_layers = bool(model_config.get("vf_share_layers"))
        self.free_log_std = bool(model_config.get("free_log_std"))
        self._params = _FlatParams(fcnet_variant_shapes(self.D, num_outputs, self.vf_share_layers, self.free_log_std))
        self._map = None                  # optional layouts: model variables -> the kernels' two-branch layout
        if self.vf_share_layers or self.free_log_std:
            self._map = fcnet_layout_map(self.D, num_outputs, self.vf_share_layers, self.free_log_std).cuda()
        if (self._map.numel() if self._map is not None else self._params.numel) != K.fcnet_num_params(self.D, self.A):
            raise DDRLError("parameter layout mismatch with libddrl_b200.so")
        gen = model_config.get("_ddrl_generator")
        flat = self._params.init_host(gen, small=("fc_out", "value_out"))
        self.theta = torch.nn.Parameter(flat.cuda()) if _HAVE_RAY else flat.cuda().requires_grad_(True)
        self.register_variables(self._params.views(self.theta.detach(), prefix=f"{name}/" if name else ""))

    def kernel_parameters(self) -> torch.Tensor:
        """The flat vector in the kernels' layout (= `theta` for the default layout), differentiable w.r.t. `theta`."""
        if self._map is None:
            return self.theta
        return torch.cat([self.theta, self.theta.new_zeros(1)])[self._map]

    def forward(self, input_dict, state, seq_lens):
        obs = input_dict["obs_flat"]
        obs = torch.as_tensor(obs, dtype=torch.float32, device=self.theta.device).contiguous()
        model_out, self._value_out = _FCNetFn.apply(self.kernel_parameters(), obs, self.A)
        return model_out, state


class FullyConnectedNetwork_GNN_GlorotUniformInitializer(_DDRLModel):
    """Separate actor GraphNet(num_outputs) and critic GraphNet(1) over Tuple(node_idx[1], obs[4,23], adj[4,4])
    (models/shared_graphnet_glorot_uniform_init.py:21-58, models/graph_net.py:10-45)."""

    def __init__(self, obs_space, action_space, num_outputs, model_config, name):
        self._setup(obs_space, action_space, num_outputs, model_config, name)
        hid = list(model_config.get("fcnet_hiddens", [64, 64]))
        if hid != [64, 64] or model_config.get("fcnet_activation", "tanh") != "tanh":
            raise DDRLError(f"{type(self).__name__}: supported model_config is fcnet_hiddens=[64,64], tanh (got {hid})")
        self_node_id_space, node_obs_space, adj_space = obs_space
        if tuple(node_obs_space.shape) != (K.GN_NODES, K.GN_FEATS + K.GN_ENC_IN):
            raise DDRLError(f"GraphNet expects node observations [4, 23] (19 leg features + 4 encoding inputs, "
                            f"models/graph_net.py:16,33-36); got {tuple(node_obs_space.shape)}")
        self.A = num_outputs // 2
        self._params = _FlatParams(graphnet_shapes(num_outputs))
        if self._params.numel != K.graphnet_num_params(num_outputs):
            raise DDRLError("parameter layout mismatch with libddrl_b200.so")
        flat = self._params.init_host(model_config.get("_ddrl_generator"), small=("actor/linear_out", "critic/linear_out"))
        self.theta = torch.nn.Parameter(flat.cuda()) if _HAVE_RAY else flat.cuda().requires_grad_(True)
        self.register_variables(self._params.views(self.theta.detach(), prefix=f"{name}/" if name else ""))

    def forward(self, input_dict, state, seq_lens):
        node_idx, obs, adj = input_dict["obs"]
        dev = self.theta.device
        node_idx = torch.as_tensor(node_idx, device=dev).reshape(-1).to(torch.int32).contiguous()
        obs = torch.as_tensor(obs, dtype=torch.float32, device=dev).contiguous()
        adj = torch.as_tensor(adj, dtype=torch.float32, device=dev).contiguous()
        action, self._value_out = _GraphNetFn.apply(self.theta, node_idx, obs, adj, self.A)
        return action, state


class _LegCouplingFn(torch.autograd.Function):
    """logits_pre [B,2A], node_idx [B], coupling [4,2] -> logits_pre * pad(coupling, ones)[node_idx]; backward returns the
    gradient w.r.t. the pre-coupling logits AND the trainable table (ddrl_leg_coupling / ddrl_leg_coupling_backward)."""

    @staticmethod
    def forward(ctx, logits_pre, node_idx, coupling):
        ctx.save_for_backward(logits_pre, node_idx, coupling)
        return K.leg_coupling_(logits_pre.detach().clone().contiguous(), node_idx, coupling.detach().contiguous())

    @staticmethod
    def backward(ctx, dout):
        logits_pre, node_idx, coupling = ctx.saved_tensors
        d = dout.contiguous().float().clone()
        dc = K.leg_coupling_backward_(d, logits_pre.detach().contiguous(), node_idx, coupling.detach().contiguous())
        return d, None, dc


class LegCoupling:
    """logits * pad(coupling[4,2], ones)[node_id]  (models/coupling_net_glorot_uniform_init.py:11-30).  `coupling` is a
    trainable variable like the reference's `tf.Variable(..., name='leg_coupling')` (:20-21)."""
    INIT = ((1.0, 1.0), (-1.0, -1.0), (-1.0, -1.0), (1.0, 1.0))

    def __init__(self, device):
        c = torch.tensor(self.INIT, dtype=torch.float32, device=device)
        self.coupling = torch.nn.Parameter(c) if _HAVE_RAY else c.requires_grad_(True)

    def __call__(self, logits_pre, node_idx):
        return _LegCouplingFn.apply(logits_pre, node_idx, self.coupling)


class FullyConnectedNetwork_Coupling_GlorotUniformInitializer(FullyConnectedNetwork_GlorotUniformInitializer):
    """FCNet over Tuple(node_idx[1], obs[D]) with the LegCoupling multiply on the logits
    (models/coupling_net_glorot_uniform_init.py:32-170).  The coupling table is applied by the ``ddrl_leg_coupling`` kernel
    and TRAINED like the reference's registered `leg_coupling` variable (:20-21,160-161): ``ddrl_leg_coupling_backward``
    returns its gradient (segment sum of dlogits * logits_pre by node id) next to the gradient of the MLP output."""

    def __init__(self, obs_space, action_space, num_outputs, model_config, name):
        orig = getattr(obs_space, "original_space", obs_space)
        self_node_id_space, leg_obs_space = orig
        if num_outputs != 4:
            raise DDRLError("LegCoupling pads coupling[4,2] with num_outputs//2 ones, which only broadcasts against "
                            "logits when num_outputs == 4 (A == 2) — same restriction as the reference")
        super().__init__(leg_obs_space, action_space, num_outputs, model_config, name)
        self.obs_space = obs_space
        self.leg_coupling = LegCoupling(self.theta.device)
        self.register_variables({f"{name}/leg_coupling": self.leg_coupling.coupling})

    def forward(self, input_dict, state, seq_lens):
        node_idx, obs = input_dict["obs"]
        dev = self.theta.device
        obs = torch.as_tensor(obs, dtype=torch.float32, device=dev).contiguous()
        node_idx = torch.as_tensor(node_idx, device=dev).reshape(-1).to(torch.int32).contiguous()
        logits, self._value_out = _FCNetFn.apply(self.kernel_parameters(), obs, self.A)
        if torch.is_grad_enabled() and (logits.requires_grad or self.leg_coupling.coupling.requires_grad):
            return self.leg_coupling(logits, node_idx), state
        return K.leg_coupling_(logits.detach().contiguous(), node_idx, self.leg_coupling.coupling.detach()), state


def register_all():
    """models/__init__.py:7-13"""
    ModelCatalog.register_custom_model("ffn", FullyConnectedNetwork_GlorotUniformInitializer)
    ModelCatalog.register_custom_model("gnn", FullyConnectedNetwork_GNN_GlorotUniformInitializer)
    ModelCatalog.register_custom_model("cup", FullyConnectedNetwork_Coupling_GlorotUniformInitializer)
    ModelCatalog.register_custom_model("fc_glorot_uniform_init", FullyConnectedNetwork_GlorotUniformInitializer)


register_all()
