"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lsqr_b200.h declares, and fails loudly (no fallback) when there is no GPU.  No compute."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from lsqrrecipes_b200 import api


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "lsqr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsqr_[a-z0-9_]+)\s*\(", text)) - {"lsqr_allreduce_max_u64_fn", "lsqr_allreduce_sum_f64_fn"})


def test_library_exports_every_declared_symbol():
    path = api.lib_path()
    assert os.path.exists(path), "liblsqr_b200.so must be built in-tree (__graft_entry__.build())"
    lib = ctypes.CDLL(path)
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/lsqr_b200.h but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == declared, "python binding and header must list the same entry points"


def test_model_table_matches_reference_shapes():
    lib = api.load_library()
    for name, m in api.MODELS.items():
        d, p, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.lsqr_model_info(m, ctypes.byref(d), ctypes.byref(p), ctypes.byref(k)) == 0
        assert (d.value, p.value, k.value) == api.MODEL_INFO[m]
    assert lib.lsqr_model_info(99, None, None, None) != 0
    # the dimension-templated estimators: ids by dimension, -1 outside the instantiated range
    for d in range(2, 9):
        assert lib.lsqr_model_plane(d) == api.MODELS[f"plane{d}"] and lib.lsqr_model_line(d) == api.MODELS[f"line{d}"]
        assert lib.lsqr_model_sphere(d) == api.MODELS["circle2" if d == 2 else f"sphere{d}"] and lib.lsqr_model_dense(d) == api.MODELS[f"dense{d}"]
    assert [f(9) for f in (lib.lsqr_model_plane, lib.lsqr_model_sphere, lib.lsqr_model_line, lib.lsqr_model_dense)] == [-1] * 4
    assert lib.lsqr_model_plane(1) == -1 and lib.lsqr_model_dense(1) == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.LsqrError):
        api.Engine("plane3", 0.5)


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", api.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under lsqrrecipes_b200/ or include/ may reference it."""
    bad = []
    for base in ("lsqrrecipes_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"pyoracle|liboracle|lsqr_oracle|libref_oracle|from oracle|import oracle", text):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_host_guards_need_no_gpu():
    """RANSAC.hxx:16-19: invalid input returns 0 and leaves `parameters` untouched (randomized overload);
    the exhaustive overload clears first (RANSAC.hxx:165-169).  Both return before any device work."""
    from lsqrrecipes_b200 import PlaneParametersEstimator, RANSAC
    est = PlaneParametersEstimator(0.5)
    assert est.numForEstimate() == 3
    params = [1.0, 2.0]
    assert RANSAC.compute(params, est, [[0, 0, 0], [1, 1, 1]], 0.99) == 0.0 and params == [1.0, 2.0]
    assert RANSAC.compute(params, est, [[0, 0, 0]] * 5, 1.0) == 0.0 and params == [1.0, 2.0]
    assert RANSAC.compute(params, est, [[0, 0, 0]] * 5, 0.0) == 0.0 and params == [1.0, 2.0]
    assert RANSAC.compute(params, est, [[0, 0, 0], [1, 1, 1]]) == 0.0 and params == []
    out = [9.0]
    est.estimate([[0, 0, 0], [1, 0, 0]], out)   # fewer than k points: cleared, nothing computed
    assert out == []


def test_estimator_mirrors_keep_the_reference_contracts():
    """Constructor contracts of the reference's estimator classes, checked without a device: the minimal subset sizes
    (`numForEstimate`, e.g. PlaneParametersEstimator.h:31 dimension, SphereParametersEstimator.h:37 dimension + 1), the
    exception for an invalid least-squares type (SphereParametersEstimator.hxx:17-18) and a clear refusal of template
    arguments that have no GPU path."""
    import lsqrrecipes_b200 as L
    want = [(L.PlaneParametersEstimator(0.5), 3), (L.PlaneParametersEstimator(0.5, dimension=4), 4),
            (L.LineParametersEstimator(0.5, dimension=2), 2), (L.LineParametersEstimator(0.5), 2), (L.Line2DParametersEstimator(0.5), 2),
            (L.SphereParametersEstimator(0.5, dimension=2), 3), (L.SphereParametersEstimator(0.5), 4),
            (L.SphereParametersEstimator(0.5, dimension=4), 5), (L.AbsoluteOrientationParametersEstimator(2.0), 3),
            (L.RayIntersectionParametersEstimator(1.0), 2), (L.PivotCalibrationEstimator(1.0), 3),
            (L.DenseLinearEquationSystemParametersEstimator(0.5, 5), 5), (L.DenseLinearEquationSystemParametersEstimator(0.5, 6), 6),
            (L.SingleUnknownPointTargetUSCalibrationParametersEstimator(1.0), 4),
            (L.CalibratedPointerTargetUSCalibrationParametersEstimator(1.0), 3)]
    for est, k in want:
        assert est.numForEstimate() == k, type(est).__name__
        assert api.MODEL_INFO[api.MODELS[est._model]][2] == k
    with pytest.raises(ValueError):
        L.SphereParametersEstimator(0.5, lsType=7)
    # the template space of the dimensioned estimators: 2..8, minimal subsets d / d + 1 / 2 / n
    for d in range(2, 9):
        assert L.PlaneParametersEstimator(0.5, dimension=d).numForEstimate() == d and L.SphereParametersEstimator(0.5, dimension=d).numForEstimate() == d + 1
        assert L.LineParametersEstimator(0.5, dimension=d).numForEstimate() == 2 and L.DenseLinearEquationSystemParametersEstimator(0.5, d).numForEstimate() == d
    for bad in (lambda: L.PlaneParametersEstimator(0.5, dimension=9), lambda: L.SphereParametersEstimator(0.5, dimension=9),
                lambda: L.LineParametersEstimator(0.5, dimension=9), lambda: L.DenseLinearEquationSystemParametersEstimator(0.5, 9)):
        with pytest.raises((NotImplementedError, ValueError)):
            bad()
