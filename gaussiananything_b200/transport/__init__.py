"""Mirror of the reference's `transport` package (/root/reference/transport/__init__.py:4-72)."""
from .transport import ModelType, PathType, Sampler, SNRType, Transport, WeightType


def create_transport(path_type='Linear', prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     snr_type='uniform'):
    """Same argument meaning and defaults as the reference factory (model prediction defaults to velocity)."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    if snr_type == "lognorm":
        snr = SNRType.LOGNORM
    elif snr_type == "uniform":
        snr = SNRType.UNIFORM
    else:
        raise ValueError(f"Invalid snr type {snr_type}")
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    if ptype == PathType.VP:
        train_eps = 1e-5 if train_eps is None else train_eps
        sample_eps = 1e-3 if sample_eps is None else sample_eps
    elif ptype in (PathType.GVP, PathType.LINEAR) and model_type != ModelType.VELOCITY:
        train_eps = 1e-3 if train_eps is None else train_eps
        sample_eps = 1e-3 if sample_eps is None else sample_eps
    else:                                   # velocity & [GVP, LINEAR] is stable everywhere
        train_eps = sample_eps = 0
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=train_eps,
                     sample_eps=sample_eps, snr_type=snr)
