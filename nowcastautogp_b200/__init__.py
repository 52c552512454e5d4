"""nowcastautogp_b200 — B200-native GP hot path behind NowcastAutoGP's public API.

Exports mirror `/root/reference/src/NowcastAutoGP.jl:10-12`. Importing the package does not load
the CUDA library; the first device call does, and raises if `libnagp.so` or a GPU is missing
(there is no CPU fallback).
"""
from .tdata import TData, create_transformed_data, create_nowcast_data
from .gpmodel import GPConfig, GPModel
from .transformations import get_transformations
from .api import make_and_fit_model, make_and_fit_models, make_and_fit_models_sharded, forecast, forecast_with_nowcasts, forecast_with_nowcasts_sharded

__all__ = ["TData", "GPModel", "GPConfig", "create_transformed_data", "make_and_fit_model", "forecast",
           "forecast_with_nowcasts", "create_nowcast_data", "forecast_with_nowcasts_sharded", "get_transformations", "make_and_fit_models", "make_and_fit_models_sharded"]
