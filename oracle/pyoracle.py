"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the two CPU checkers.

  Oracle("port")  -> oracle/liboracle.so          (plain-C restatement, lsqr_oracle.c)
  Oracle("ref")   -> oracle/_ref/libref_oracle.so (the reference's own sources + VNL shim)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under lsqrrecipes_b200/ does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i32p = ctypes.POINTER(ctypes.c_int32)

MODELS = {"plane3": 0, "line2d": 1, "line2": 2, "line3": 3, "circle2": 4, "sphere3": 5, "absor": 6, "ray": 7, "pivot": 8,
          "dense5": 9, "dense6": 10, "usxw": 11, "uscp": 12, "sphere4": 13, "plane4": 14}
# model -> (D doubles per datum, P params, k minimal subset)
INFO = {0: (3, 6, 3), 1: (2, 4, 2), 2: (2, 4, 2), 3: (3, 6, 2), 4: (2, 3, 3), 5: (3, 4, 4), 6: (6, 7, 3), 7: (6, 3, 2), 8: (12, 6, 3),
        9: (6, 5, 5), 10: (7, 6, 6), 11: (14, 20, 4), 12: (17, 17, 3), 13: (4, 5, 5), 14: (4, 8, 4)}
# the rest of the reference's template space: PlaneParametersEstimator<2, 5..8>, SphereParametersEstimator<5..8>,
# LineParametersEstimator<4..8>, DenseLinearEquationSystemParametersEstimator<double, 2..4, 7, 8> (ids as in lsqr_oracle.c)
for _d, _id in ((2, 15), (5, 16), (6, 17), (7, 18), (8, 19)):
    MODELS[f"plane{_d}"] = _id
    INFO[_id] = (_d, 2 * _d, _d)
for _d in range(5, 9):
    MODELS[f"sphere{_d}"] = 20 + _d - 5
    INFO[20 + _d - 5] = (_d, _d + 1, _d + 1)
for _d in range(4, 9):
    MODELS[f"line{_d}"] = 24 + _d - 4
    INFO[24 + _d - 4] = (_d, 2 * _d, 2)
for _n, _id in ((2, 29), (3, 30), (4, 31), (7, 32), (8, 33)):
    MODELS[f"dense{_n}"] = _id
    INFO[_id] = (_n + 1, _n, _n)


def lib_path(kind):
    return os.path.join(_HERE, "liboracle.so") if kind == "port" else os.path.join(_HERE, "_ref", "libref_oracle.so")


def available(kind):
    return os.path.exists(lib_path(kind))


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        self.pfx = "orc_" if kind == "port" else "ref_"
        self.lib = ctypes.CDLL(lib_path(kind))
        f = self._fn
        f("estimate", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp, ctypes.c_size_t, _dp])
        f("least_squares", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp, ctypes.c_size_t, _dp])
        f("agree", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp, ctypes.c_int, _dp, ctypes.c_size_t, _u8p])
        f("score_subsets", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp, ctypes.c_size_t, _i32p, ctypes.c_size_t, _u32p, _dp, ctypes.c_int])
        f("num_threads", ctypes.c_int, [])
        f("model_info", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)])
        f("weighted_absor", ctypes.c_int, [_dp, ctypes.c_size_t, _dp, _dp])
        if kind == "port":
            f("ransac_exhaustive", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp, ctypes.c_size_t, _dp, _u8p, _dp, _u32p, ctypes.POINTER(ctypes.c_uint64)])
            f("choose", ctypes.c_uint, [ctypes.c_uint, ctypes.c_uint])
            f("unrank_lex", None, [ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, _i32p])
            f("num_tries", ctypes.c_uint, [ctypes.c_double, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint])
            f("last_lm", None, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)])
        else:
            f("ransac", ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, _dp, ctypes.c_size_t, ctypes.c_int, ctypes.c_double, _dp, _u8p, _dp])

    def model_info(self, model):
        """(D, P, k) as the C side has them, or None for an unknown id."""
        d, p, k = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if self._model_info(model, ctypes.byref(d), ctypes.byref(p), ctypes.byref(k)) != 0:
            return None
        return d.value, p.value, k.value

    def _fn(self, name, res, args):
        fn = getattr(self.lib, self.pfx + name)
        fn.restype = res
        fn.argtypes = args
        setattr(self, "_" + name, fn)

    @staticmethod
    def _data(model, data):
        d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, INFO[model][0])
        return d

    def num_threads(self):
        return self._num_threads()

    def estimate(self, model, delta, data, aux=0.0):
        d = self._data(model, data)
        out = np.zeros(24)
        n = self._estimate(model, delta, aux, _ptr(d, _dp), d.shape[0], _ptr(out, _dp))
        return out[:max(n, 0)].copy()

    def least_squares(self, model, delta, data, ls_type=1, aux=0.0):
        d = self._data(model, data)
        out = np.zeros(24)
        n = self._least_squares(model, delta, aux, ls_type, _ptr(d, _dp), d.shape[0], _ptr(out, _dp))
        return out[:max(n, 0)].copy()

    def weighted_absor(self, data, weights):
        """AbsoluteOrientationParametersEstimator::weightedLeastSquaresEstimate"""
        d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, 6)
        w = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
        assert len(w) == d.shape[0]
        out = np.zeros(24)
        n = self._weighted_absor(_ptr(d, _dp), d.shape[0], _ptr(w, _dp), _ptr(out, _dp))
        return out[:max(n, 0)].copy()

    def agree(self, model, delta, params, data, aux=0.0):
        d = self._data(model, data)
        p = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(d.shape[0], dtype=np.uint8)
        c = self._agree(model, delta, aux, _ptr(p, _dp), len(p), _ptr(d, _dp), d.shape[0], _ptr(out, _u8p))
        return c, out

    def score_subsets(self, model, delta, data, subsets, aux=0.0, nthreads=0, want_params=True):
        d = self._data(model, data)
        _, P, k = INFO[model]
        s = np.ascontiguousarray(subsets, dtype=np.int32).reshape(-1, k)
        counts = np.zeros(s.shape[0], dtype=np.uint32)
        params = np.zeros((s.shape[0], P)) if want_params else None
        rc = self._score_subsets(model, delta, aux, _ptr(d, _dp), d.shape[0], _ptr(s, _i32p), s.shape[0], _ptr(counts, _u32p), _ptr(params, _dp), nthreads)
        assert rc == 0
        return counts, params

    def ransac_exhaustive(self, model, delta, data, ls_type=1, aux=0.0):
        d = self._data(model, data)
        out = np.zeros(24)
        mask = np.zeros(d.shape[0], dtype=np.uint8)
        frac = ctypes.c_double(0)
        if self.kind == "port":
            bc = ctypes.c_uint32(0)
            br = ctypes.c_uint64(0)
            n = self._ransac_exhaustive(model, delta, aux, ls_type, _ptr(d, _dp), d.shape[0], _ptr(out, _dp), _ptr(mask, _u8p), ctypes.byref(frac), ctypes.byref(bc), ctypes.byref(br))
            return out[:max(n, 0)].copy(), mask, frac.value, bc.value, br.value
        n = self._ransac(model, delta, aux, ls_type, _ptr(d, _dp), d.shape[0], 1, 0.0, _ptr(out, _dp), _ptr(mask, _u8p), ctypes.byref(frac))
        return out[:max(n, 0)].copy(), mask, frac.value, int(mask.sum()), None

    def ransac_random(self, model, delta, data, prob, ls_type=1, aux=0.0):
        assert self.kind == "ref"
        d = self._data(model, data)
        out = np.zeros(24)
        mask = np.zeros(d.shape[0], dtype=np.uint8)
        frac = ctypes.c_double(0)
        n = self._ransac(model, delta, aux, ls_type, _ptr(d, _dp), d.shape[0], 0, prob, _ptr(out, _dp), _ptr(mask, _u8p), ctypes.byref(frac))
        return out[:max(n, 0)].copy(), mask, frac.value

    # port-only helpers
    def last_lm(self):
        """(MINPACK info, function evaluations) of the last Levenberg-Marquardt refit."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self._last_lm(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def choose(self, n, m):
        return self._choose(n, m)

    def unrank_lex(self, rank, n, k):
        out = np.zeros(k, dtype=np.int32)
        self._unrank_lex(rank, n, k, _ptr(out, _i32p))
        return out

    def num_tries(self, prob, votes, n, k, all_tries):
        return self._num_tries(prob, votes, n, k, all_tries)
