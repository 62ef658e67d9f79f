"""The C++ drop-in headers (include/lsqrRecipes/*.h) are the boundary a user of the reference sees.
CPU: they compile and link against the in-tree C-ABI library.  GPU: the demo that drives every
estimator through RANSAC<T,S>::compute passes all of its checks."""
import os
import subprocess

import pytest

from conftest import ROOT

EX = os.path.join(ROOT, "examples")


def _build():
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    subprocess.check_call(["make", "-C", EX], env=env, stdout=subprocess.DEVNULL)
    return os.path.join(EX, "dropin_demo")


def test_dropin_headers_compile_and_link():
    exe = _build()
    assert os.path.exists(exe)
    # every reference header of the path has a same-named counterpart
    want = ["ParametersEstimator.h", "RANSAC.h", "PlaneParametersEstimator.h", "LineParametersEstimator.h", "Line2DParametersEstimator.h",
            "SphereParametersEstimator.h", "AbsoluteOrientationParametersEstimator.h", "RayIntersectionParametersEstimator.h",
            "PivotCalibrationParametersEstimator.h", "Point.h", "Point2D.h", "Point3D.h", "Vector.h", "Vector3D.h", "Frame.h", "Ray3D.h", "Epsilon.h"]
    have = set(os.listdir(os.path.join(ROOT, "include", "lsqrRecipes")))
    assert not [h for h in want if h not in have]


def test_dropin_demo_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([_build()], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_dropin_demo_on_gpu():
    r = subprocess.run([_build()], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL CHECKS PASSED" in r.stdout
