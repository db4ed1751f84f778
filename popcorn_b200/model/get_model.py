"""Model registry with the reference's surface (model/get_model.py:9-60): Args, model_dict,
calculate_input_channels, get_model_kwargs — same names, argument meaning and error behaviour."""
from __future__ import annotations

from typing import Any, Dict, NamedTuple

from .popcorn import POPCORN


class Args(NamedTuple):
    Sentinel1: bool
    NIR: bool
    Sentinel2: bool
    feature_extractor: str
    occupancymodel: bool
    pretrained: bool
    biasinit: float
    sentinelbuildings: bool


model_dict = {"POPCORN": POPCORN}


def calculate_input_channels(args) -> int:
    """2 for Sentinel-1 (VV,VH), 3 for Sentinel-2 RGB, +1 for NIR (model/get_model.py:23-32)."""
    return (2 if args.Sentinel1 else 0) + (1 if args.NIR else 0) + (3 if args.Sentinel2 else 0)


def get_model_kwargs(args, model_name: str) -> Dict[str, Any]:
    """Constructor kwargs for model_dict[model_name] (model/get_model.py:34-60)."""
    if model_name not in model_dict:
        raise ValueError(f"Model {model_name} not found in model dictionary")
    return {
        "input_channels": calculate_input_channels(args),
        "feature_extractor": args.feature_extractor,
        "occupancymodel": args.occupancymodel,
        "pretrained": args.pretrained,
        "biasinit": args.biasinit,
        "sentinelbuildings": args.sentinelbuildings,
    }
