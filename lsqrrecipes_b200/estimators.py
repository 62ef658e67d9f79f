"""Python mirror of the reference's operator interface for this path, over the C ABI.

Same names, argument meaning and error behaviour as the C++ classes
(parametersEstimators/ParametersEstimator.h:26-64 and the concrete estimators), so that the
parity tests read like the reference's own tests.  The authoritative drop-in for C++ callers
is include/lsqrRecipes/*.h; both sit on liblsqr_b200.so and neither computes on the host.

  * ``estimate(data, parameters)`` / ``leastSquaresEstimate(data, parameters)`` clear and fill the
    ``parameters`` list; an empty list is the reference's "degenerate / failed" signal.
  * ``agree(parameters, datum)`` returns bool.
  * ``RANSAC.compute(parameters, estimator, data, desiredProbabilityForNoOutliers, consensusSet)``
    and the 4-argument exhaustive overload return the inlier fraction (RANSAC.h:75-79,111-113).
"""
import numpy as np

from . import api


class ParametersEstimator:
    """ParametersEstimator<T,S> (ParametersEstimator.h:26-64)."""

    _model = None

    def __init__(self, min_elements, delta, aux=0.0, ls_type=api.LS_GEOMETRIC):
        self.minForEstimate = min_elements
        self._delta, self._aux, self._ls_type = float(delta), float(aux), int(ls_type)
        self._engine = None

    def numForEstimate(self):
        return self.minForEstimate

    # the engine is created on first use so that constructing an estimator needs no GPU
    def engine(self):
        if self._engine is None:
            self._engine = api.Engine(self._model, self._delta, self._aux, self._ls_type)
        return self._engine

    def _reconfigure(self):
        if self._engine is not None:
            self._engine.set_estimator(self._delta, self._aux, self._ls_type)

    def setDelta(self, delta):
        self._delta = float(delta)
        self._reconfigure()

    def estimate(self, data, parameters):
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, api.MODEL_INFO[api.MODELS[self._model]][0])
        if d.shape[0] < self.minForEstimate:
            return
        parameters.extend(self.engine().estimate(d).tolist())

    def leastSquaresEstimate(self, data, parameters):
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, api.MODEL_INFO[api.MODELS[self._model]][0])
        parameters.extend(self.engine().least_squares(d).tolist())

    def agree(self, parameters, datum):
        return bool(self.engine().agree(np.asarray(parameters, dtype=np.float64), np.asarray(datum, dtype=np.float64))[0])


class PlaneParametersEstimator(ParametersEstimator):
    """PlaneParametersEstimator<dimension> for dimension 2..8 (PlaneParametersEstimator.h:24-91)."""

    def __init__(self, delta, dimension=3):
        if not 2 <= dimension <= 8:
            raise NotImplementedError("hyperplanes are accelerated for d = 2..8")
        self._model = f"plane{dimension}"
        super().__init__(dimension, delta)


class LineParametersEstimator(ParametersEstimator):
    """LineParametersEstimator<dimension> for dimension 2..8 (LineParametersEstimator.h:35-100)."""

    def __init__(self, delta, dimension=3):
        if not 2 <= dimension <= 8:
            raise NotImplementedError("kD lines are accelerated for d = 2..8")
        self._model = f"line{dimension}"
        super().__init__(2, delta)


class Line2DParametersEstimator(ParametersEstimator):
    """Line2DParametersEstimator (Line2DParametersEstimator.h:22-85); parameters [n_x,n_y,a_x,a_y]."""
    _model = "line2d"

    def __init__(self, delta):
        super().__init__(2, delta)


class SphereParametersEstimator(ParametersEstimator):
    """SphereParametersEstimator<dimension> for dimension 2 (circle) .. 8 (SphereParametersEstimator.h:29-190)."""
    ALGEBRAIC, GEOMETRIC = api.LS_ALGEBRAIC, api.LS_GEOMETRIC

    def __init__(self, delta, lsType=api.LS_GEOMETRIC, dimension=3):
        if lsType not in (self.ALGEBRAIC, self.GEOMETRIC):
            raise ValueError("lsType must be ALGEBRAIC or GEOMETRIC")  # SphereParametersEstimator.hxx:17-18 throws
        if not 2 <= dimension <= 8:
            raise NotImplementedError("hyperspheres are accelerated for d = 2..8")
        self._model = "circle2" if dimension == 2 else f"sphere{dimension}"
        super().__init__(dimension + 1, delta, ls_type=lsType)

    def setLeastSquaresType(self, lsType):
        self._ls_type = int(lsType)
        self._reconfigure()


class AbsoluteOrientationParametersEstimator(ParametersEstimator):
    """AbsoluteOrientationParametersEstimator; datum = (p_first, p_second) as 6 doubles; parameters [s,qx,qy,qz,tx,ty,tz]."""
    _model = "absor"

    def __init__(self, delta):
        super().__init__(3, delta)

    def weightedLeastSquaresEstimate(self, data, weights, parameters):
        """AbsoluteOrientationParametersEstimator.cxx:208-297: Horn's method with one weight per pair."""
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, 6)
        if d.shape[0] < self.minForEstimate:
            return
        parameters.extend(self.engine().weighted_least_squares(d, weights).tolist())


class RayIntersectionParametersEstimator(ParametersEstimator):
    """RayIntersectionParametersEstimator; datum = Ray3D (p, n) as 6 doubles; parameters [x,y,z]."""
    _model = "ray"

    def __init__(self, delta, minimalAngularDeviation=0.017453292519943295769236907684886):
        super().__init__(2, delta, aux=minimalAngularDeviation)

    def leastSquaresEstimate(self, data, parameters):
        # RayIntersectionParametersEstimator.cxx:100-144 neither clears `parameters` nor checks the size
        d = np.asarray(data, dtype=np.float64).reshape(-1, 6)
        parameters.extend(self.engine().least_squares(d).tolist())


class PivotCalibrationEstimator(ParametersEstimator):
    """PivotCalibrationEstimator; datum = Frame as 12 doubles (R row-major, t); parameters [tDRF, tW]."""
    _model = "pivot"

    def __init__(self, delta):
        super().__init__(3, delta)

    def leastSquaresEstimate(self, data, parameters):
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, 12)
        if d.shape[0] < self.minForEstimate:
            return
        parameters.extend(self.engine().least_squares(d).tolist())


class DenseLinearEquationSystemParametersEstimator(ParametersEstimator):
    """DenseLinearEquationSystemParametersEstimator<double, n> (n = 2..8); datum = AugmentedRow as n+1 doubles
    [a_0..a_{n-1}, b]; parameters = the solution x (DenseLinearEquationSystemParametersEstimator.hxx:17-119)."""

    def __init__(self, delta, n):
        if not 2 <= n <= 8:
            raise ValueError("dense linear systems are instantiated for n = 2..8 (the reference exercises 5 and 6)")
        self._model = f"dense{n}"
        super().__init__(n, delta)


class SingleUnknownPointTargetUSCalibrationParametersEstimator(ParametersEstimator):
    """SingleUnknownPointTargetUSCalibrationParametersEstimator (cross-wire phantom ultrasound calibration,
    SinglePointTargetUSCalibrationParametersEstimator.cxx:10-329); datum = (Frame T2, Point2D q) as 14 doubles
    [R2 row-major, t2, u, v]; 20 parameters [t1, t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]."""
    _model = "usxw"
    ANALYTIC, ITERATIVE = 0, 1

    def __init__(self, delta, lsType=1):
        super().__init__(4, delta, ls_type=lsType)

    def setLeastSquaresType(self, lsType):
        self._ls_type = int(lsType)
        self._reconfigure()

    def estimate(self, data, parameters):
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, 14)
        if d.shape[0] != self.minForEstimate:      # exactly four (.cxx:21-22)
            return
        parameters.extend(self.engine().estimate(d).tolist())


class CalibratedPointerTargetUSCalibrationParametersEstimator(ParametersEstimator):
    """CalibratedPointerTargetUSCalibrationParametersEstimator (SinglePointTargetUSCalibrationParametersEstimator.cxx:663-985);
    datum = (Frame T2, Point2D q, Point3D p) as 17 doubles [R2 row-major, t2, u, v, p]; 17 parameters
    [t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]."""
    _model = "uscp"
    ANALYTIC, ITERATIVE = 0, 1

    def __init__(self, delta, lsType=1):
        super().__init__(3, delta, ls_type=lsType)

    def setLeastSquaresType(self, lsType):
        self._ls_type = int(lsType)
        self._reconfigure()

    def estimate(self, data, parameters):
        parameters.clear()
        d = np.asarray(data, dtype=np.float64).reshape(-1, 17)
        if d.shape[0] != self.minForEstimate:      # exactly three (.cxx:674-675)
            return
        parameters.extend(self.engine().estimate(d).tolist())


class RANSAC:
    """RANSAC<T,S> (RANSAC.h:47-151): two static compute() overloads."""

    precision = api.FP32   # fast consensus; the consensus set and the refine are always fp64
    seed = 0

    @staticmethod
    def compute(parameters, paramEstimator, data, desiredProbabilityForNoOutliers=None, consensusSet=None):
        exhaustive = desiredProbabilityForNoOutliers is None
        dim = api.MODEL_INFO[api.MODELS[paramEstimator._model]][0]
        d = np.asarray(data, dtype=np.float64).reshape(-1, dim)
        n, k = d.shape[0], paramEstimator.numForEstimate()
        if exhaustive:
            parameters.clear()                      # RANSAC.hxx:165 clears before the size check
            if n < k:
                return 0.0
        else:
            p = desiredProbabilityForNoOutliers
            if n < k or p >= 1.0 or p <= 0.0:       # RANSAC.hxx:16-19 returns BEFORE parameters.clear() (:43)
                return 0.0
            parameters.clear()
        eng = paramEstimator.engine()
        eng.upload(d)
        want_mask = consensusSet is not None
        if exhaustive:
            r = eng.ransac_exhaustive(precision=api.FP64, want_mask=want_mask)
        else:
            r = eng.ransac(desiredProbabilityForNoOutliers, precision=RANSAC.precision, seed=RANSAC.seed, want_mask=want_mask)
        if r["best_count"] > 0:
            parameters.extend(r["params"].tolist())
            if want_mask:
                consensusSet.clear()
                consensusSet.extend(bool(b) for b in r["mask"])
        return r["fraction"]
