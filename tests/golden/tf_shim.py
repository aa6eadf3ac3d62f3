"""A numpy stand-in for the handful of TensorFlow / Keras / RLlib names the reference's model files use — so that
`/root/reference/models/*.py` can be IMPORTED AND EXECUTED UNMODIFIED in a container without TensorFlow or Ray
(tests/golden/make_models_golden.py).  Test infrastructure for generating golden vectors; nothing in the product or in
the test suite imports it.

What this is and is not: every op below follows the documented TensorFlow semantics of the op of the same name (argument
order, axes, index order of `tf.where`, empty segments of `unsorted_segment_mean` = 0, `leaky_relu` alpha 0.2, VarianceScaling
limits) — but it is numpy, not TensorFlow.  Vectors made with it pin the reference's COMPOSITION of these ops (which op,
on which axis, in which order, with which weights), i.e. everything the reference authors wrote; they cannot pin TensorFlow's
own kernels.  Tensors are plain `numpy.ndarray`s (so `@`, `**`, slicing and `+=` behave as in eager TF); float64 inputs keep
float64 throughout.

Keras is emulated in eager form plus a minimal functional API: a layer called on a symbolic `KTensor` records a graph node and
immediately evaluates a one-row zero probe through it (that is how `build(input_shape)` learns its shapes);
`keras.Model(inputs, outputs)(x)` replays the graph on real arrays."""
from __future__ import annotations

import sys
import types

import numpy as np

float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64
newaxis = None


# ---- eager ops -----------------------------------------------------------------------------------------------------------
def _a(x):
    return x.value if isinstance(x, KTensor) else np.asarray(x)


def Variable(initial_value, dtype=None, trainable=True, name=None):
    return np.array(initial_value, dtype=dtype)


def zeros(shape, dtype=float32):
    return np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)


def ones_like(x):
    return np.ones_like(x)


def shape(x, out_type=int32):
    return np.asarray(np.shape(x), dtype=out_type)


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def reshape(x, shp):
    return np.reshape(x, [int(s) for s in np.atleast_1d(shp)])


def reduce_sum(x, axis=None):
    return np.sum(x, axis=axis)


def reduce_prod(x, axis=None):
    return np.prod(x, axis=axis)


def exp(x):
    return np.exp(x)


def tanh(x):
    return np.tanh(x)


def minimum(a, b):
    return np.minimum(a, b)


def eye(n):
    return np.eye(int(n), dtype=np.float32)


def concat(values, axis):
    return np.concatenate(list(values), axis=axis)


def stack(values, axis=0):
    return np.stack(list(values), axis=axis)


def unstack(x, axis=0):
    x = np.asarray(x)
    return [np.take(x, i, axis=axis) for i in range(x.shape[axis])]


def expand_dims(x, axis):
    return np.expand_dims(x, axis)


def squeeze(x, axis=None):
    return np.squeeze(x, axis=axis)


def tile(x, multiples):
    return np.tile(x, [int(m) for m in multiples])


def pad(x, paddings, constant_values=0):
    return np.pad(x, [tuple(int(v) for v in p) for p in paddings], constant_values=constant_values)


def where(condition):
    """Single-argument tf.where: coordinates of the true / non-zero elements, row-major order, int64 [n, rank]."""
    return np.argwhere(np.asarray(condition)).astype(np.int64)


def gather(params, indices, axis=0, batch_dims=0):
    params, indices = np.asarray(params), np.asarray(indices)
    if batch_dims == 0:
        return np.take(params, indices, axis=axis)
    if batch_dims == 1 and axis == 1 and indices.ndim == 1:
        return params[np.arange(params.shape[0]), indices]
    raise NotImplementedError("tf_shim.gather: only batch_dims=0, or batch_dims=1 with axis=1 and 1-D indices")


def gather_nd(params, indices):
    indices = np.asarray(indices)
    return np.asarray(params)[tuple(indices[:, i] for i in range(indices.shape[1]))]


def scatter_nd(indices, updates, shp):
    out = np.zeros([int(s) for s in shp], dtype=np.asarray(updates).dtype)
    np.add.at(out, tuple(np.asarray(indices)[:, i] for i in range(np.asarray(indices).shape[1])), updates)
    return out


def _segment_sum(data, segment_ids, num_segments):
    data = np.asarray(data)
    out = np.zeros((int(num_segments),) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, np.asarray(segment_ids), data)
    return out


def _segment_mean(data, segment_ids, num_segments):
    data = np.asarray(data)
    s = _segment_sum(data, segment_ids, num_segments)
    n = np.zeros(int(num_segments), dtype=data.dtype)
    np.add.at(n, np.asarray(segment_ids), 1)
    n = n.reshape((-1,) + (1,) * (data.ndim - 1))
    return np.where(n > 0, s / np.maximum(n, 1), 0).astype(data.dtype)          # empty segment -> 0


def _diag(v):
    v = np.asarray(v)
    return v[..., :, None] * np.eye(v.shape[-1], dtype=v.dtype)


def _leaky_relu(x, alpha=0.2):
    return np.where(x >= 0, x, alpha * x)


# ---- Keras ---------------------------------------------------------------------------------------------------------------
class VarianceScaling:
    """tf.keras.initializers.VarianceScaling: uniform -> U(-l, l), l = sqrt(3 * scale / n), n = fan_avg | fan_in | fan_out."""
    rng = np.random.default_rng(0)

    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        self.scale, self.mode, self.distribution, self.seed = scale, mode, distribution, seed

    def __call__(self, shp, dtype=np.float32):
        fan_in, fan_out = shp
        n = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode]
        if self.distribution != "uniform":
            raise NotImplementedError(self.distribution)
        lim = np.sqrt(3.0 * self.scale / n)
        return VarianceScaling.rng.uniform(-lim, lim, size=shp).astype(dtype)


class KTensor:
    """Symbolic tensor of the functional API: op + parents, and the value of a one-row zero probe."""

    def __init__(self, fn, parents, value, layer=None):
        self.fn, self.parents, self.value, self.layer = fn, parents, value, layer

    @property
    def shape(self):
        return (None,) + tuple(self.value.shape[1:])


def _evaluate(t, feed, memo):
    if not isinstance(t, KTensor):
        return t
    if id(t) in feed:
        return feed[id(t)]
    if id(t) not in memo:
        memo[id(t)] = t.fn(*[_evaluate(p, feed, memo) for p in t.parents])
    return memo[id(t)]


def Input(shape=None, dtype=None, name=None):
    t = KTensor(None, [], np.zeros((1,) + tuple(shape), dtype=dtype or np.float32))
    t.name = name
    return t


class Layer:
    _registry: list = []          # creation order, for Model.layers

    def __init__(self, name=None, **kwargs):
        self.name, self.built = name, False
        Layer._registry.append(self)

    def build(self, input_shape):
        pass

    def call(self, *args):
        raise NotImplementedError

    def _run(self, *args):
        if not self.built:
            self.build(tuple(np.shape(args[0])))
            self.built = True
        return self.call(*args)

    def __call__(self, *args):
        flat = [a for arg in args for a in (arg if isinstance(arg, (list, tuple)) else [arg])]
        if any(isinstance(a, KTensor) for a in flat):
            parents = list(flat)
            shapes = [len(arg) if isinstance(arg, (list, tuple)) else None for arg in args]

            def fn(*vals, self=self, shapes=shapes):
                it, rebuilt = iter(vals), []
                for n in shapes:
                    rebuilt.append(next(it) if n is None else [next(it) for _ in range(n)])
                return self._run(*rebuilt)
            probe = fn(*[_a(p) for p in parents])
            return KTensor(fn, parents, probe, layer=self)
        return self._run(*args)

    @property
    def variables(self):
        out = []
        for v in vars(self).values():
            if isinstance(v, np.ndarray):
                out.append(v)
            elif isinstance(v, Layer):
                out.extend(v.variables)
        return out


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, name=None, **kwargs):
        super().__init__(name=name)
        self.units, self.use_bias = int(units), use_bias
        self.activation = Activation(activation)
        self.kernel_initializer = kernel_initializer or VarianceScaling(1.0, "fan_avg", "uniform")      # glorot_uniform
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        self.kernel = self.kernel_initializer((int(input_shape[-1]), self.units))
        if self.use_bias:
            self.bias = np.zeros(self.units, dtype=np.float32)

    def call(self, x):
        y = np.asarray(x) @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return self.activation.call(y)


class Activation(Layer):
    def __init__(self, activation=None, **kwargs):
        super().__init__()
        self.fn = {None: None, "linear": None, "tanh": np.tanh, "relu": lambda v: np.maximum(v, 0)}.get(activation, activation) \
            if (activation is None or isinstance(activation, str)) else activation

    def call(self, x):
        return x if self.fn is None else self.fn(x)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__()
        self.function = function

    def call(self, x):
        return self.function(x)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__()
        self.axis = axis

    def call(self, xs):
        return np.concatenate(list(xs), axis=self.axis)


class Model(Layer):
    """Subclassed use: `super().__init__()` then `call`; functional use: `Model(inputs, outputs)`."""

    def __init__(self, inputs=None, outputs=None, **kwargs):
        super().__init__()
        self.inputs, self.outputs = inputs, outputs
        self.built = True

    def __call__(self, *args):
        if self.outputs is None:
            return self.call(*args)
        ins = self.inputs if isinstance(self.inputs, (list, tuple)) else [self.inputs]
        vals = args[0] if isinstance(self.inputs, (list, tuple)) else [args[0]]
        feed, memo = {id(t): np.asarray(v) for t, v in zip(ins, vals)}, {}
        outs = [_evaluate(t, feed, memo) for t in (self.outputs if isinstance(self.outputs, (list, tuple)) else [self.outputs])]
        return outs if isinstance(self.outputs, (list, tuple)) else outs[0]

    @property
    def layers(self):
        """Layers reachable from the outputs, in creation order."""
        seen, stack_ = set(), list(self.outputs if isinstance(self.outputs, (list, tuple)) else [self.outputs])
        found = []
        while stack_:
            t = stack_.pop()
            if isinstance(t, KTensor) and id(t) not in seen:
                seen.add(id(t))
                if t.layer is not None:
                    found.append(t.layer)
                stack_.extend(t.parents)
        return [l for l in Layer._registry if any(l is f for f in found)]

    def get_layer(self, name):
        return next(l for l in self.layers if l.name == name)

    @property
    def variables(self):
        if self.outputs is None:
            return Layer.variables.fget(self)
        return [v for l in self.layers for v in l.variables]


# ---- RLlib ---------------------------------------------------------------------------------------------------------------
class TFModelV2:
    def __init__(self, obs_space, action_space, num_outputs, model_config, name):
        self.obs_space, self.action_space, self.num_outputs = obs_space, action_space, num_outputs
        self.model_config, self.name, self.var_list = model_config, name, []

    def register_variables(self, variables):
        self.var_list.extend(variables)

    def variables(self):
        return list(self.var_list)


class _Catalog:
    registered: dict = {}

    @staticmethod
    def register_custom_model(name, cls):
        _Catalog.registered[name] = cls


def get_activation_fn(name, framework="tf"):
    return {None: None, "linear": None, "tanh": np.tanh, "relu": (lambda v: np.maximum(v, 0))}[name]


def install():
    """Put the stand-ins into sys.modules under the names the reference imports."""
    this = sys.modules[__name__]
    tf = types.ModuleType("tensorflow")
    for k in ("float32", "float64", "int32", "int64", "newaxis", "Variable", "zeros", "ones_like", "shape", "cast", "reshape",
              "reduce_sum", "reduce_prod", "exp", "tanh", "minimum", "eye", "concat", "stack", "unstack", "expand_dims",
              "squeeze", "tile", "pad", "where", "gather", "gather_nd", "scatter_nd"):
        setattr(tf, k, getattr(this, k))
    tf.math = types.SimpleNamespace(unsorted_segment_sum=_segment_sum, unsorted_segment_mean=_segment_mean)
    tf.linalg = types.SimpleNamespace(diag=_diag)
    tf.nn = types.SimpleNamespace(tanh=np.tanh, leaky_relu=_leaky_relu)
    layers = types.ModuleType("tensorflow.keras.layers")
    for k in ("Layer", "Dense", "Activation", "Lambda", "Concatenate", "Input"):
        setattr(layers, k, getattr(this, k))
    initializers = types.ModuleType("tensorflow.keras.initializers")
    initializers.VarianceScaling = VarianceScaling
    keras = types.ModuleType("tensorflow.keras")
    keras.layers, keras.initializers, keras.Model = layers, initializers, Model
    tf.keras = keras
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.initializers": initializers}
    ray = types.ModuleType("ray")
    rllib = types.ModuleType("ray.rllib")
    models = types.ModuleType("ray.rllib.models")
    models.ModelCatalog = _Catalog
    mtf = types.ModuleType("ray.rllib.models.tf")
    mv2 = types.ModuleType("ray.rllib.models.tf.tf_modelv2")
    mv2.TFModelV2 = TFModelV2
    utils = types.ModuleType("ray.rllib.utils")
    fw = types.ModuleType("ray.rllib.utils.framework")
    fw.get_activation_fn = get_activation_fn
    fw.try_import_tf = lambda: (tf, tf, 2)
    ray.rllib, rllib.models, rllib.utils, models.tf, mtf.tf_modelv2, utils.framework = rllib, models, utils, mtf, mv2, fw
    mods.update({"ray": ray, "ray.rllib": rllib, "ray.rllib.models": models, "ray.rllib.models.tf": mtf,
                 "ray.rllib.models.tf.tf_modelv2": mv2, "ray.rllib.utils": utils, "ray.rllib.utils.framework": fw})
    sys.modules.update(mods)
    return tf
