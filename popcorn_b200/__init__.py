"""popcorn_b200 — B200-native (sm_100a) implementation of POPCORN's dense-prediction hot path.

Public surface mirrors the reference (model/get_model.py, model/popcorn.py):
    from popcorn_b200.model.get_model import model_dict, get_model_kwargs, Args
    model = model_dict["POPCORN"](**get_model_kwargs(args, "POPCORN")).cuda()
    out = model(sample, padding=False)          # {"popcount", "popdensemap", "scale"}
The compute lives in popcorn_b200/libpopcorn_b200.so (C-ABI: include/popcorn_b200.h).
"""
from . import _lib, ops, weights  # noqa: F401
from .model import POPCORN, Args, get_model_kwargs, model_dict  # noqa: F401

__version__ = "0.1.0"
