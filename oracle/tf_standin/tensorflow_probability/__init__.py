"""Stand-in for the two tensorflow_probability entry points the reference touches (test infrastructure; see
../tensorflow/__init__.py).  Categorical.sample is served from the scripted RNG: mfp.py:34-43,301."""
import sys
import types

import torch

from tensorflow import config


class Categorical:
    def __init__(self, logits=None, probs=None):
        self.logits = torch.as_tensor(logits)
        self.support = [i for i, v in enumerate(self.logits.tolist()) if v > float("-inf")]

    def sample(self, n):
        out = config.rng.categorical(int(n)).to(torch.int32)
        assert all(int(v) in self.support for v in out), "scripted task id outside the distribution's support"
        return out


class Bernoulli:
    def __init__(self, probs=None):
        self.probs = probs

    def sample(self, shape):
        raise NotImplementedError("unused_masking is not on the MFP hot path")


distributions = types.ModuleType(__name__ + ".distributions")
distributions.Categorical = Categorical
distributions.Bernoulli = Bernoulli
sys.modules[__name__ + ".distributions"] = distributions


class MultivariateNormalDiag:  # import-time alias in models/canvasvae.py:14 (VAE baselines: not on the hot path)
    def __init__(self, *a, **k):
        raise NotImplementedError("VAE baselines are not on the MFP hot path")


distributions.MultivariateNormalDiag = MultivariateNormalDiag
