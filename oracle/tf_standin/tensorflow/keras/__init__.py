"""Keras slice of the TensorFlow stand-in (test infrastructure; see ../__init__.py for what it pins).

Layer bookkeeping follows what the reference relies on: sub-layers held in attributes, dicts and lists are tracked
(Keras wraps those containers), variables are named by their attribute path (the TF object-graph path, which is also
what oracle/mfp_oracle.py::variable_specs uses), ``add_loss`` / ``add_metric`` / regulariser losses are collected
at the root, and Dense / Embedding build lazily.  Appendix-A semantics are marked [recall].
"""
import math as _math
import sys
import types

import torch

from .. import config


def _mod(name, **members):
    m = types.ModuleType(name)
    m.__dict__.update(members)
    sys.modules[name] = m
    return m


# ------------------------------------------------------------------------------------------------ regularizers
class L2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def __call__(self, w):
        return self.l2 * (w * w).sum()  # [recall] A3: l2 * sum(w^2), no 1/2


regularizers = _mod(__name__ + ".regularizers", l2=L2, L2=L2)


# ------------------------------------------------------------------------------------------------ layers
class Layer:
    def __init__(self, name=None, **kwargs):
        assert not kwargs or set(kwargs) <= {"dtype", "trainable"}, kwargs
        self.name = name
        self._losses, self._metrics, self._own = [], [], {}
        self.supports_masking = False

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    # -- tracking
    def _children(self):
        for attr, val in list(self.__dict__.items()):
            if attr.startswith("_"):
                continue
            if isinstance(val, Layer):
                yield attr, val
            elif isinstance(val, dict):
                for k, v in val.items():
                    if isinstance(v, Layer):
                        yield "%s/%s" % (attr, k), v
            elif isinstance(val, (list, tuple)):
                for i, v in enumerate(val):
                    if isinstance(v, Layer):
                        yield "%s/%d" % (attr, i), v

    def named_variables(self, prefix=""):
        out = {}
        for k, v in self._own.items():
            out[prefix + k] = v
        for path, child in self._children():
            out.update(child.named_variables(prefix + path + "/"))
        return out

    def add_weight(self, name, shape, init, regularizer=None):
        from .. import Tensor

        w = _initialise(shape, init).as_subclass(Tensor).requires_grad_(True)
        self._own[name] = w
        if regularizer is not None:
            self.__dict__.setdefault("_regularized", []).append((name, regularizer))
        return w

    def _all_layers(self):
        yield self
        for _, c in self._children():
            yield from c._all_layers()

    def add_loss(self, value):
        self._losses.append(value)

    def add_metric(self, value, name=None):
        self._metrics.append((name, value))

    def reset_step_state(self):
        for layer in self._all_layers():
            layer._losses, layer._metrics = [], []

    @property
    def losses(self):
        """add_loss values of this call plus every regulariser term ([recall] A2: Keras adds both to the loss)."""
        out = []
        for layer in self._all_layers():
            out.extend(layer._losses)
            for name, reg in layer.__dict__.get("_regularized", []):
                out.append(reg(layer._own[name]))
        return out

    @property
    def step_metrics(self):
        out = []
        for layer in self._all_layers():
            out.extend(layer._metrics)
        return out


def _initialise(shape, init):
    """Placeholders only: the golden generator overwrites every variable with the oracle's init_params values."""
    if init == "ones":
        return torch.ones(shape, dtype=config.FLOAT)
    return torch.zeros(shape, dtype=config.FLOAT)


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, kernel_regularizer=None, bias_regularizer=None, **kw):
        super().__init__(name=name)
        assert not kw, kw
        self.units, self.activation = int(units), activation
        self._kreg, self._breg = kernel_regularizer, bias_regularizer

    def build(self, in_dim):
        self.add_weight("kernel", (in_dim, self.units), "glorot", self._kreg)
        self.add_weight("bias", (self.units,), "zeros", self._breg)

    def call(self, x, training=None):
        if "kernel" not in self._own:
            self.build(int(x.shape[-1]))
        y = x.to(config.FLOAT) @ self._own["kernel"] + self._own["bias"]  # [recall] A12
        if self.activation == "relu":
            y = torch.relu(y)
        else:
            assert self.activation is None
        return y


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, name=None, embeddings_regularizer=None, **kw):
        super().__init__(name=name)
        assert not kw, kw
        self.add_weight("embeddings", (int(input_dim), int(output_dim)), "uniform", embeddings_regularizer)

    def call(self, ids, training=None):
        ids = torch.as_tensor(ids)
        if ids.is_floating_point():  # Keras casts non-integer ids to int32 (encoder.py:167-172 passes float zeros/ones)
            ids = ids.to(torch.int32)
        return self._own["embeddings"][ids.to(torch.int64)]


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, name=None, **kw):  # [recall] A1: default epsilon 1e-3, last axis
        super().__init__(name=name)
        assert not kw, kw
        self.epsilon = epsilon

    def call(self, x, training=None):
        if "gamma" not in self._own:
            self.add_weight("gamma", (int(x.shape[-1]),), "ones")
            self.add_weight("beta", (int(x.shape[-1]),), "zeros")
        mu = x.mean(dim=-1, keepdim=True)
        var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
        return (x - mu) / torch.sqrt(var + self.epsilon) * self._own["gamma"] + self._own["beta"]


class Dropout(Layer):
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name)
        self.rate = float(rate)

    def call(self, x, training=None):
        if not training or self.rate == 0.0:
            return x
        keep = config.rng.dropout(tuple(x.shape)).to(x.dtype)
        return x * keep * (1.0 / (1.0 - self.rate))  # [recall] A10: inverted dropout


class ReLU(Layer):
    def call(self, x):
        return torch.relu(x)


class Activation(Layer):
    def __init__(self, kind, **kw):
        super().__init__(**kw)
        assert kind == "relu"

    def call(self, x):
        return torch.relu(x)


class GlobalAveragePooling1D(Layer):
    def call(self, x, mask=None):
        raise NotImplementedError("not on the MFP hot path")


class Sequential(Layer):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self._seq = list(layers or [])
        # TF object-graph names of a Sequential's children: layer_with_weights-<i>
        for i, layer in enumerate(self._seq):
            setattr(self, "layer_with_weights-%d" % i, layer)

    def call(self, x, training=None):
        for layer in self._seq:
            x = layer(x)
        return x


class Model(Layer):
    def compile(self, optimizer=None, run_eagerly=None, **kw):
        self.optimizer = optimizer

    def train_step_standin(self, inputs):
        """[recall] Keras default train_step with loss=None: forward(training=True); loss = sum(self.losses);
        gradients w.r.t. trainable variables; optimizer.apply_gradients."""
        self.reset_step_state()
        outputs = self(inputs, training=True)
        loss = sum(self.losses)
        names = list(self.named_variables().keys())
        variables = [self.named_variables()[n] for n in names]
        grads = torch.autograd.grad(loss, variables, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(v) for g, v in zip(grads, variables)]
        if getattr(self, "optimizer", None) is not None:
            self.optimizer.apply_gradients(zip(grads, variables))
        return outputs, loss.detach(), dict(zip(names, grads))


_exp = _mod(__name__ + ".layers.experimental")
_pre = _mod(__name__ + ".layers.experimental.preprocessing",
            Discretization=type("Discretization", (Layer,), {}), StringLookup=type("StringLookup", (Layer,), {}),
            IntegerLookup=type("IntegerLookup", (Layer,), {}))
_exp.preprocessing = _pre
layers = _mod(__name__ + ".layers", Layer=Layer, Dense=Dense, Embedding=Embedding, LayerNormalization=LayerNormalization,
              Dropout=Dropout, ReLU=ReLU, Activation=Activation, GlobalAveragePooling1D=GlobalAveragePooling1D,
              Sequential=Sequential, experimental=_exp)


# ------------------------------------------------------------------------------------------------ losses / metrics
_EPS = 1e-7  # [recall] A6: keras.backend.epsilon()


def sparse_categorical_crossentropy(y_true, y_pred, from_logits=False):
    """[recall] A6 (eager, probabilities in): clip to [eps, 1-eps], log, then softmax-CE *on the logs*."""
    assert not from_logits
    z = torch.log(torch.clamp(y_pred, _EPS, 1.0 - _EPS))
    picked = torch.gather(z, -1, y_true.to(torch.int64)[..., None])[..., 0]
    return torch.logsumexp(z, dim=-1) - picked


def mean_squared_error(y_true, y_pred):
    return ((y_pred - y_true.to(y_pred.dtype)) ** 2).mean(dim=-1)  # [recall] A7


def _l2_normalize(x):
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim=-1, keepdim=True), min=1e-12))  # [recall] A8


def cosine_similarity(y_true, y_pred, axis=-1):
    return -(_l2_normalize(y_true.to(y_pred.dtype)) * _l2_normalize(y_pred)).sum(dim=-1)  # Keras returns the NEGATIVE cosine


losses = _mod(__name__ + ".losses", sparse_categorical_crossentropy=sparse_categorical_crossentropy,
              mean_squared_error=mean_squared_error, cosine_similarity=cosine_similarity)
metrics = _mod(__name__ + ".metrics", mean_absolute_error=lambda a, b: (a - b).abs().mean(dim=-1))


# ------------------------------------------------------------------------------------------------ optimizers
class Adam:
    """[recall] A4/A5: per-variable clip_by_norm, then TF Adam with epsilon outside the bias-corrected sqrt."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=None):
        self.lr, self.b1, self.b2, self.eps, self.clipnorm = learning_rate, beta_1, beta_2, epsilon, clipnorm
        self.t, self.m, self.v = 0, {}, {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        alpha = self.lr * _math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for g, w in grads_and_vars:
                if self.clipnorm is not None:
                    n = torch.sqrt((g * g).sum())
                    g = g * (self.clipnorm / torch.clamp(n, min=self.clipnorm))
                k = id(w)
                m = self.m.get(k, torch.zeros_like(w))
                v = self.v.get(k, torch.zeros_like(w))
                self.m[k] = m = self.b1 * m + (1.0 - self.b1) * g
                self.v[k] = v = self.b2 * v + (1.0 - self.b2) * g * g
                w.sub_(alpha * m / (torch.sqrt(v) + self.eps))  # explicit in-place: the Tensor subclass makes `-=` rebind


optimizers = _mod(__name__ + ".optimizers", Adam=Adam)
