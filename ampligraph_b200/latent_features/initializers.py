"""Initialiser selection of the facade: what `tf.keras.initializers.get` resolves for
`entity_relation_initializer` in the reference (EmbeddingLookupLayer.set_initializer,
layers/encoding/EmbeddingLookupLayer.py:105-129; default 'glorot_uniform',
models/ScoringBasedEmbeddingModel.py:1149).

Every named initialiser is reduced to one of the four device kinds of `kge_init_table`
(uniform / normal / truncated_normal / constant, include/kge_b200.h) on the table shape
`[rows, internal_k]`, with Keras' conventions for a 2-D weight: fan_in = rows, fan_out = internal_k.
The random stream is Philox4x32-10 keyed by the model seed (TensorFlow's own stream is not
reproducible outside TF, SURVEY.md 8c: the reference's tests pin mean/std only,
tests/ampligraph/latent_features/test_initializers.py:48-49).
"""
import math

_TRUNC_STD_FIX = 0.87962566103423978  # stddev of N(0,1) truncated to [-2, 2] (Keras VarianceScaling)


class Initializer:
    """Base class: `spec(rows, cols)` -> (kind, a, b) for kge_init_table."""
    name = None

    def spec(self, rows, cols):
        raise NotImplementedError

    def get_config(self):
        return {}


class RandomUniform(Initializer):
    name = "random_uniform"

    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = float(minval), float(maxval), seed

    def spec(self, rows, cols):
        return "uniform", self.minval, self.maxval

    def get_config(self):
        return {"minval": self.minval, "maxval": self.maxval, "seed": self.seed}


class RandomNormal(Initializer):
    name = "random_normal"

    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = float(mean), float(stddev), seed

    def spec(self, rows, cols):
        return "normal", self.mean, self.stddev

    def get_config(self):
        return {"mean": self.mean, "stddev": self.stddev, "seed": self.seed}


class TruncatedNormal(RandomNormal):
    name = "truncated_normal"

    def spec(self, rows, cols):
        return "truncated_normal", self.mean, self.stddev


class Constant(Initializer):
    name = "constant"

    def __init__(self, value=0.0):
        self.value = float(value)

    def spec(self, rows, cols):
        return "constant", self.value, 0.0

    def get_config(self):
        return {"value": self.value}


class Zeros(Constant):
    name = "zeros"

    def __init__(self):
        super().__init__(0.0)


class Ones(Constant):
    name = "ones"

    def __init__(self):
        super().__init__(1.0)


class VarianceScaling(Initializer):
    """tf.keras.initializers.VarianceScaling on a [rows, cols] weight (fan_in = rows, fan_out = cols)."""
    name = "variance_scaling"

    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        if scale <= 0.0:
            raise ValueError("`scale` must be positive float. Received: scale=%r" % (scale,))
        if mode not in ("fan_in", "fan_out", "fan_avg"):
            raise ValueError("Invalid `mode` argument: %r" % (mode,))
        distribution = {"normal": "truncated_normal"}.get(distribution, distribution)
        if distribution not in ("uniform", "truncated_normal", "untruncated_normal"):
            raise ValueError("Invalid `distribution` argument: %r" % (distribution,))
        self.scale, self.mode, self.distribution, self.seed = float(scale), mode, distribution, seed

    def spec(self, rows, cols):
        fan = {"fan_in": max(1.0, rows), "fan_out": max(1.0, cols), "fan_avg": max(1.0, (rows + cols) / 2.0)}[self.mode]
        scale = self.scale / fan
        if self.distribution == "uniform":
            limit = math.sqrt(3.0 * scale)
            return "uniform", -limit, limit
        if self.distribution == "truncated_normal":
            return "truncated_normal", 0.0, math.sqrt(scale) / _TRUNC_STD_FIX
        return "normal", 0.0, math.sqrt(scale)

    def get_config(self):
        return {"scale": self.scale, "mode": self.mode, "distribution": self.distribution, "seed": self.seed}


def _vs(name, scale, mode, distribution):
    def make(seed=None):
        v = VarianceScaling(scale, mode, distribution, seed)
        v.name = name
        return v
    return make


GlorotUniform = _vs("glorot_uniform", 1.0, "fan_avg", "uniform")   # U(+-sqrt(6/(rows+cols)))
GlorotNormal = _vs("glorot_normal", 1.0, "fan_avg", "truncated_normal")
HeUniform = _vs("he_uniform", 2.0, "fan_in", "uniform")
HeNormal = _vs("he_normal", 2.0, "fan_in", "truncated_normal")
LecunUniform = _vs("lecun_uniform", 1.0, "fan_in", "uniform")
LecunNormal = _vs("lecun_normal", 1.0, "fan_in", "truncated_normal")

_BY_NAME = {
    "random_uniform": RandomUniform, "uniform": RandomUniform, "randomuniform": RandomUniform,
    "random_normal": RandomNormal, "normal": RandomNormal, "randomnormal": RandomNormal,
    "truncated_normal": TruncatedNormal, "truncatednormal": TruncatedNormal,
    "zeros": Zeros, "ones": Ones, "constant": Constant,
    "variance_scaling": VarianceScaling, "variancescaling": VarianceScaling,
    "glorot_uniform": GlorotUniform, "glorotuniform": GlorotUniform, "xavier_uniform": GlorotUniform,
    "glorot_normal": GlorotNormal, "glorotnormal": GlorotNormal, "xavier_normal": GlorotNormal,
    "he_uniform": HeUniform, "heuniform": HeUniform, "he_normal": HeNormal, "henormal": HeNormal,
    "lecun_uniform": LecunUniform, "lecununiform": LecunUniform, "lecun_normal": LecunNormal, "lecunnormal": LecunNormal,
}


def get(identifier):
    """tf.keras.initializers.get: name | {'class_name':, 'config':} | Initializer instance | callable(shape) | array.
    Arrays and callables are returned unchanged (the model packs their values into the table)."""
    if identifier is None:
        return GlorotUniform()
    if isinstance(identifier, Initializer):
        return identifier
    if isinstance(identifier, str):
        key = identifier.lower()
        if key not in _BY_NAME:
            raise ValueError("Could not interpret initializer identifier: %r" % (identifier,))
        return _BY_NAME[key]()
    if isinstance(identifier, dict):
        key = str(identifier.get("class_name", "")).lower()
        if key not in _BY_NAME:
            raise ValueError("Could not interpret initializer identifier: %r" % (identifier,))
        return _BY_NAME[key](**dict(identifier.get("config", {})))
    if callable(identifier) or hasattr(identifier, "shape") or isinstance(identifier, (list, tuple)):
        return identifier
    raise ValueError("Could not interpret initializer identifier: %r" % (identifier,))
