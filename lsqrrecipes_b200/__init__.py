"""lsqrrecipes_b200 -- B200-native RANSAC / least-squares engine behind the
ParametersEstimator / RANSAC::compute API of zivy/LSQRRecipes.

The product is the C-ABI shared library ``liblsqr_b200.so`` (hand-written sm_100a CUDA,
``include/lsqr_b200.h``) plus the re-authored drop-in C++ headers in ``include/lsqrRecipes/``.
This Python package is a thin ctypes binding over the same C ABI, used by the tests, by
``bench.py`` and for torch.distributed plumbing.  It never computes anything itself and it
has no CPU fallback: importing works without a GPU, every call fails loudly without one.
"""
from .api import (  # noqa: F401
    Engine,
    LsqrError,
    MODELS,
    MODEL_INFO,
    FP32,
    FP64,
    SAMPLE_PHILOX,
    SAMPLE_EXHAUSTIVE,
    SAMPLE_LIST,
    SAMPLE_PARAMS,
    LS_ALGEBRAIC,
    LS_GEOMETRIC,
    lib_path,
    load_library,
)
from .estimators import (  # noqa: F401
    RANSAC,
    PlaneParametersEstimator,
    LineParametersEstimator,
    Line2DParametersEstimator,
    SphereParametersEstimator,
    AbsoluteOrientationParametersEstimator,
    RayIntersectionParametersEstimator,
    PivotCalibrationEstimator,
    DenseLinearEquationSystemParametersEstimator,
    SingleUnknownPointTargetUSCalibrationParametersEstimator,
    CalibratedPointerTargetUSCalibrationParametersEstimator,
)

__all__ = [
    "Engine", "LsqrError", "MODELS", "MODEL_INFO", "FP32", "FP64", "RANSAC",
    "PlaneParametersEstimator", "LineParametersEstimator", "Line2DParametersEstimator",
    "SphereParametersEstimator", "AbsoluteOrientationParametersEstimator",
    "RayIntersectionParametersEstimator", "PivotCalibrationEstimator", "DenseLinearEquationSystemParametersEstimator", "SingleUnknownPointTargetUSCalibrationParametersEstimator", "CalibratedPointerTargetUSCalibrationParametersEstimator",
]
