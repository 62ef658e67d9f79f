"""Levenberg-Marquardt, pinned on the CPU (no GPU needed).

The reference's nonlinear refits go through vnl_levenberg_marquardt = MINPACK lmder (VNL / netlib absent from the image).
Three links of the chain are checked here:
  1. the restatement of lmder in oracle/minpack_lm.h against the REAL MINPACK (scipy.optimize.leastsq wraps netlib's lmder):
     same info code, same number of function evaluations, same end point, for the reference's sphere and cross-wire functors
     with the reference's tolerances;
  2. the engine's on-device controller (lsqrrecipes_b200/csrc/lm_minpack.cuh, plain C++ when g++ compiles it) driven by the
     moments J^T J, J^T f, |f|^2 of each evaluation, against that lmder on the full Jacobian: same MINPACK info / outcome and
     end points within the 1e-6 the north star asks of converged Levenberg-Marquardt results;
  3. the golden fixture of the reference's own experimental cross-wire data set against the port.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from lsqrrecipes_b200 import synth
from oracle.pyoracle import MODELS

_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lmchk(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lmchk") / "liblmchk.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-fopenmp", "-shared", "-w", "-o", out,
                           os.path.join(ROOT, "tests", "lm_controller_check.cxx")])
    lib = ctypes.CDLL(out)
    lib.lmchk_run.argtypes = [ctypes.c_int, _dp, ctypes.c_int, _dp, _dp]
    lib.lmchk_run.restype = ctypes.c_int
    lib.lmchk_eval.argtypes = [ctypes.c_int, _dp, ctypes.c_int, _dp, ctypes.c_int, _dp]

    def run(model, data, init):
        d = np.ascontiguousarray(data, dtype=np.float64)
        x0 = np.ascontiguousarray(init, dtype=np.float64)
        out_ = np.zeros(48)
        n = lib.lmchk_run(model, d.ctypes.data_as(_dp), d.shape[0], x0.ctypes.data_as(_dp), out_.ctypes.data_as(_dp))
        return {"info": int(out_[0]), "nfev": int(out_[1]), "ctrl_info": int(out_[2]), "ctrl_nfev": int(out_[3]), "ctrl_status": int(out_[4]),
                "x": out_[6:6 + n].copy(), "ctrl_x": out_[6 + n:6 + 2 * n].copy()}

    def functor(model, data, nparams):
        """f and J of the oracle's restatement of the reference's functor, as Python callables (bit-identical values)."""
        d = np.ascontiguousarray(data, dtype=np.float64)

        def f(x):
            out_ = np.zeros(d.shape[0])
            lib.lmchk_eval(model, d.ctypes.data_as(_dp), d.shape[0], np.ascontiguousarray(x).ctypes.data_as(_dp), 1, out_.ctypes.data_as(_dp))
            return out_

        def jac(x):
            out_ = np.zeros((d.shape[0], nparams))
            lib.lmchk_eval(model, d.ctypes.data_as(_dp), d.shape[0], np.ascontiguousarray(x).ctypes.data_as(_dp), 2, out_.ctypes.data_as(_dp))
            return out_
        return f, jac
    run.functor = functor
    return run


def _inliers(port, name, n, seed):
    m, delta = MODELS[name], synth.DELTAS[name]
    data, true = synth.GENERATORS[name](n, seed=seed)
    _, flags = port.agree(m, delta, true, data)
    inl = np.ascontiguousarray(data[flags.astype(bool)])
    init = port.least_squares(m, delta, inl, 0)
    return inl, (init[:8] if name == "uscp" else init[:11])


def test_lmder_restatement_matches_real_minpack(port, lmchk):
    """oracle/minpack_lm.h vs netlib's lmder (through scipy.optimize.leastsq, mode 1, factor 100) on bit-identical functors:
    same stopping reason, same number of function evaluations, same end point to the last bits.  Sphere with the reference's
    settings (ftol 1e-10, xtol = gtol = 1e-15, 500 evaluations); cross-wire and calibrated pointer on synthetic inliers; and
    the reference's own cross-wire data file with all tolerances 1e-15 / 5000 evaluations (136 evaluations of a flat valley)."""
    leastsq = pytest.importorskip("scipy.optimize").leastsq
    cases = []
    for name, n, nprm, tol in (("sphere3", 3000, 4, (1e-10, 1e-15, 1e-15, 500)), ("circle2", 500, 3, (1e-10, 1e-15, 1e-15, 500)),
                               ("usxw", 300, 11, (1e-15, 1e-15, 1e-15, 5000)), ("uscp", 300, 8, (1e-7, 1e-7, 1e-7, 5000))):
        inl, init = _inliers(port, name, n, seed=5)
        cases.append((name, inl, init, nprm, tol))
    g = golden("usxw_file")
    cases.append(("usxw", g["data"], g["ls0"][:11], 11, (1e-15, 1e-15, 1e-15, 5000)))
    for name, inl, init, nprm, (ftol, xtol, gtol, maxfev) in cases:
        f, jac = lmchk.functor(MODELS[name], inl, nprm)
        x, _, info, _, ier = leastsq(f, np.array(init, dtype=np.float64), Dfun=jac, full_output=True, ftol=ftol, xtol=xtol, gtol=gtol, maxfev=maxfev, factor=100)
        r = lmchk(MODELS[name], inl, init)
        assert (r["info"], r["nfev"]) == (ier, info["nfev"]), name
        assert np.allclose(r["x"], x, rtol=1e-13, atol=1e-13), name
    # the valley is flat: the reference's own functor rounds differently from the port's and ends 1e-8 away
    assert np.allclose(r["x"], g["ls1"][:11], rtol=1e-6, atol=1e-6), "fixture = what the reference (oracle/_ref) returned for its data file"


@pytest.mark.parametrize("name,n", [("circle2", 3000), ("sphere3", 300), ("sphere3", 20000), ("sphere4", 3000), ("usxw", 300), ("usxw", 3000), ("uscp", 300), ("uscp", 20000)])
def test_device_controller_follows_lmder(port, lmchk, name, n):
    """The engine's controller works from J^T J, J^T f and |f|^2 (what one streaming pass delivers); lmder from the m x n
    Jacobian.  Same outcome, same end point."""
    inl, init = _inliers(port, name, n, seed=5)
    r = lmchk(MODELS[name], inl, init)
    assert (r["info"] in (1, 2, 3, 4)) == (r["ctrl_status"] == 1)
    if name != "usxw":       # quadratically convergent cases: identical stopping reason; identical evaluation count where steps are taken
        assert r["ctrl_info"] == r["info"]
        if name != "uscp":
            assert r["ctrl_nfev"] == r["nfev"]
    assert np.allclose(r["ctrl_x"], r["x"], rtol=1e-6, atol=1e-6)
    # the calibrated-pointer fit stops at the analytic start (scaled gradient / step below the reference's 1e-7 tolerances)
    if name == "uscp":
        assert np.array_equal(r["x"], init)


def test_crosswire_file_fixture(port):
    """The reference's experimental cross-wire data (testing/Data/crossWirePhantom*.txt, 54 poses): analytic and iterative
    least squares of the port against what the reference returned."""
    g = golden("usxw_file")
    m = MODELS["usxw"]
    assert g["data"].shape == (54, 14) and len(g["ls1"]) == 20
    assert np.allclose(port.least_squares(m, 5.0, g["data"], 0), g["ls0"], rtol=1e-9, atol=1e-9)
    assert np.allclose(port.least_squares(m, 5.0, g["data"], 1), g["ls1"], rtol=1e-6, atol=1e-6)
    assert np.abs(g["ls1"][:11] - g["ls0"][:11]).max() > 1e-3     # the iteration does move away from the analytic start
