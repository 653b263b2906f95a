"""Knobs of the TensorFlow stand-in (test infrastructure; see tensorflow/__init__.py)."""
import torch

FLOAT = torch.float32  # what tf.float32 maps to


class ScriptedRNG:
    """Every tf.random.* / tfp sample / Dropout mask request is served from a caller-provided script, in call order.
    Script items: ("uniform"|"randint"|"normal"|"categorical"|"dropout", array or None).  ``None`` = the draw is never
    used by the reference (e.g. the eagerly-evaluated "random" entry of apply_token's dict, masking.py:80-93) and is
    answered with a constant."""

    def __init__(self, script=None):
        self.script = list(script or [])
        self.pos = 0
        self.log = []

    def _next(self, kind, shape):
        self.log.append((kind, tuple(shape)))
        if self.pos >= len(self.script):
            raise RuntimeError("RNG script exhausted at call %d: %s %s" % (self.pos, kind, shape))
        want, arr = self.script[self.pos]
        self.pos += 1
        if want != kind:
            raise RuntimeError("RNG script mismatch at call %d: reference asked %s%s, script has %s" % (self.pos - 1, kind, shape, want))
        if arr is not None and tuple(arr.shape) != tuple(shape):
            raise RuntimeError("RNG script shape mismatch at call %d (%s): %s vs %s" % (self.pos - 1, kind, shape, arr.shape))
        return arr

    def uniform(self, shape):
        a = self._next("uniform", shape)
        return torch.full(shape, 0.5) if a is None else torch.as_tensor(a)

    def randint(self, shape, lo, hi):
        a = self._next("randint", shape)
        return torch.full(shape, lo, dtype=torch.int32) if a is None else torch.as_tensor(a)

    def normal(self, shape, stddev):
        a = self._next("normal", shape)  # scripted values already include the stddev
        return torch.zeros(shape) if a is None else torch.as_tensor(a)

    def categorical(self, n):
        return torch.as_tensor(self._next("categorical", (n,)))

    def dropout(self, shape):
        return torch.as_tensor(self._next("dropout", shape))

    def done(self):
        return self.pos == len(self.script)


rng = ScriptedRNG()
