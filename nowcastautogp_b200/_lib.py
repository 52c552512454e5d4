"""ctypes binding of libnagp.so (include/nagp.h). There is no CPU fallback: a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAGP_LIB", os.path.join(_HERE, "libnagp.so"))

E_ARG, E_CUDA, E_PROGRAM, E_SIZE = -1, -2, -3, -4

_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double

# name -> (restype, argtypes); mirrors include/nagp.h one to one
SIGNATURES = {
    "nagp_version": (_i32, []),
    "nagp_init": (_i32, [_i32, C.POINTER(_vp)]),
    "nagp_destroy": (None, [_vp]),
    "nagp_last_error": (C.c_char_p, [_vp]),
    "nagp_set_stream": (_i32, [_vp, _vp]),
    "nagp_set_jitter": (_i32, [_vp, _f64]),
    "nagp_launch_count": (_i64, [_vp]),
    "nagp_set_variant": (_i32, [_vp, _i32]),
    "nagp_last_kernel": (_i32, [_vp]),
    "nagp_logml_batch": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _f64, _vp, _i64, _vp, _vp]),
    "nagp_forecast_instances": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _f64,
                                       _i64, _i64, _i64, _vp, _vp, _f64, _vp, _vp, _f64, _f64, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "nagp_factor_store": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _f64, _i64, _i64, _i64, _vp, _vp, _f64,
                                 _vp, _f64, _f64, _vp, C.POINTER(_vp), _vp, _vp]),
    "nagp_factor_free": (None, [_vp]),
    "nagp_logml_grad": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _f64,
                               _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nagp_hmc": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f64, _f64, _vp, _vp, _i64, _i64, _vp, _vp,
                        _f64, _vp, _i64, _vp, _i64, _i64, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nagp_factor_store_large": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _f64, _vp,
                                       C.POINTER(_vp), _vp, _vp]),
    "nagp_factor_append": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nagp_factor_size": (_i64, [_vp]),
    "nagp_append": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "nagp_predict": (_i32, [_vp, _vp, _vp, _vp]),
    "nagp_ess": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "nagp_draw": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _f64,
                         _vp, _vp, _vp, _vp]),
    "nagp_forecast_with_nowcasts": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _f64,
                                           _i64, _i64, _i64, _vp, _vp, _f64, _vp, _vp, _f64, _f64, _vp,
                                           _vp, _vp, _vp, _f64, _vp, _vp, _vp, _vp, _vp]),
    "nagp_forecast_with_nowcasts_theta": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _f64,
                                                 _i64, _i64, _i64, _vp, _vp, _f64, _vp, _vp, _f64, _f64, _vp,
                                                 _vp, _vp, _vp, _f64, _vp, _vp, _vp, _vp, _vp]),
    "nagp_forecast_summary": (_i32, [_vp, _i32, _f64, _f64, _f64, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libnagp.so; raises if it has not been built (python -m nowcastautogp_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback). "
                "Build it with `python -m nowcastautogp_b200.build`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
