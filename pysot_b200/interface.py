"""Host-side mirror of the reference's surrogate plugin interface.

When pySOT itself is importable the GPU surrogate subclasses ``pySOT.surrogate.RBFInterpolant`` (so that
``isinstance(s, pySOT.surrogate.Surrogate)`` holds, as ``pySOT/strategy/surrogate_strategy.py:154`` asserts)
and takes pySOT's own kernel / tail objects.  When it is not (the GPU box of this build has no pySOT), the
structurally identical classes below provide the same names, attributes and error behaviour:

  Kernel / CubicKernel / TPSKernel / LinearKernel      pySOT/surrogate/kernels/*.py   (order, eval, deriv)
  Tail / LinearTail / ConstantTail                     pySOT/surrogate/tails/*.py     (degree, dim, dim_tail, eval, deriv)
  identity / median_capping                            pySOT/surrogate/output_transformations/*.py
  Surrogate                                            pySOT/surrogate/surrogate.py:20-98

Kernel and tail objects are descriptors for the GPU path (they select the device code by class name); their
numpy ``eval`` / ``deriv`` exist because they are public API of the reference, not because the surrogate
calls them.
"""
import abc

import numpy as np

try:  # pragma: no cover - exercised only where pySOT is installed
    import pySOT.surrogate as _ps

    HAVE_PYSOT = True
except Exception:  # ImportError, or pySOT's own optional-dependency failures
    _ps = None
    HAVE_PYSOT = False

_EPS = np.finfo(float).eps


class _Kernel(abc.ABC):
    order = None

    @abc.abstractmethod
    def eval(self, dists):
        ...

    @abc.abstractmethod
    def deriv(self, dists):
        ...


class _CubicKernel(_Kernel):
    """phi(r) = r^3, conditionally positive definite of order 2 (kernels/cubic_kernel.py)."""

    order = 2

    def eval(self, dists):
        return np.asarray(dists) ** 3

    def deriv(self, dists):
        return 3 * np.asarray(dists) ** 2


class _TPSKernel(_Kernel):
    """phi(r) = r^2 log r, order 2; r < eps is clamped to eps in place (kernels/tps_kernel.py:28,41)."""

    order = 2

    def eval(self, dists):
        dists[dists < _EPS] = _EPS
        return (dists ** 2) * np.log(dists)

    def deriv(self, dists):
        dists[dists < _EPS] = _EPS
        return dists * (1 + 2 * np.log(dists))


class _LinearKernel(_Kernel):
    """phi(r) = r, order 1 (kernels/linear_kernel.py)."""

    order = 1

    def eval(self, dists):
        return dists

    def deriv(self, dists):
        return np.ones(np.shape(dists))


class _Tail(abc.ABC):
    degree = None

    @abc.abstractmethod
    def eval(self, X):
        ...

    @abc.abstractmethod
    def deriv(self, x):
        ...

    def _check(self, X):
        X = np.atleast_2d(X)
        if X.shape[1] != self.dim:
            raise ValueError("Input has the wrong number of dimensions")
        return X


class _LinearTail(_Tail):
    """basis {1, x_1, .., x_d} (tails/linear_tail.py)."""

    degree = 1

    def __init__(self, dim):
        self.dim = dim
        self.dim_tail = 1 + dim

    def eval(self, X):
        X = self._check(X)
        return np.hstack((np.ones((X.shape[0], 1)), X))

    def deriv(self, x):
        x = self._check(x)
        return np.hstack((np.zeros((self.dim, 1)), np.eye(self.dim)))


class _ConstantTail(_Tail):
    """basis {1} (tails/constant_tail.py)."""

    degree = 0

    def __init__(self, dim):
        self.dim = dim
        self.dim_tail = 1

    def eval(self, X):
        X = self._check(X)
        return np.ones((X.shape[0], 1))

    def deriv(self, x):
        x = self._check(x)
        return np.zeros((self.dim, 1))


def _identity(x):
    return x


def _median_capping(x):
    """values above the median are replaced by the median (output_transformations/median_capping.py)."""
    x = x.copy()
    med = np.median(x)
    x[x > med] = med
    return x


def _to_unit_box(x, lb, ub):
    return (x - lb) / (ub - lb)  # pySOT/utils.py:33


class _Surrogate(abc.ABC):
    """Point bookkeeping of pySOT/surrogate/surrogate.py:33-74."""

    def __init__(self, dim, lb, ub, output_transformation=None):
        self.dim = dim
        self.lb = lb
        self.ub = ub
        self.output_transformation = _identity if output_transformation is None else output_transformation
        self.reset()

    def reset(self):
        self.num_pts = 0
        self.X = np.empty([0, self.dim])
        self._X = np.empty([0, self.dim])
        self.fX = np.empty([0, 1])
        self.updated = False

    def add_points(self, xx, fx):
        """Append evaluations; never fits (surrogate.py:50-74).  fx may be a float, 0-d, 1-d or (k, 1)."""
        xx = np.atleast_2d(xx)
        if isinstance(fx, float):
            fx = np.array([fx])
        fx = np.asarray(fx)
        if fx.ndim == 0:
            fx = np.expand_dims(fx, axis=0)
        if fx.ndim == 1:
            fx = np.expand_dims(fx, axis=1)
        assert xx.shape[0] == fx.shape[0] and xx.shape[1] == self.dim
        self.X = np.vstack((self.X, xx))
        self._X = _to_unit_box(self.X, self.lb, self.ub)  # every row is re-mapped, as in the reference (:71)
        self.fX = np.vstack((self.fX, fx))
        self.num_pts += xx.shape[0]
        self.updated = False

    @abc.abstractmethod
    def predict(self, xx):
        ...

    @abc.abstractmethod
    def predict_deriv(self, xx):
        ...


if HAVE_PYSOT:  # pragma: no cover
    Kernel, CubicKernel, TPSKernel, LinearKernel = _ps.Kernel, _ps.CubicKernel, _ps.TPSKernel, _ps.LinearKernel
    Tail, LinearTail, ConstantTail = _ps.Tail, _ps.LinearTail, _ps.ConstantTail
    identity, median_capping = _ps.identity, _ps.median_capping
    Surrogate = _ps.Surrogate
    ReferenceRBFInterpolant = _ps.RBFInterpolant
else:
    Kernel, CubicKernel, TPSKernel, LinearKernel = _Kernel, _CubicKernel, _TPSKernel, _LinearKernel
    Tail, LinearTail, ConstantTail = _Tail, _LinearTail, _ConstantTail
    identity, median_capping = _identity, _median_capping
    Surrogate = _Surrogate
    ReferenceRBFInterpolant = None

KERNEL_IDS = {"CubicKernel": 0, "TPSKernel": 1, "LinearKernel": 2, "_CubicKernel": 0, "_TPSKernel": 1, "_LinearKernel": 2}
TAIL_IDS = {"LinearTail": 0, "ConstantTail": 1, "_LinearTail": 0, "_ConstantTail": 1}


def kernel_id(kernel):
    """Device kernel selector; anything but the reference's three kernels has no GPU code and no fallback."""
    try:
        return KERNEL_IDS[type(kernel).__name__]
    except KeyError:
        raise NotImplementedError("no sm_100a code for kernel %r (supported: Cubic, TPS, Linear)" % type(kernel).__name__)


def tail_id(tail):
    try:
        return TAIL_IDS[type(tail).__name__]
    except KeyError:
        raise NotImplementedError("no sm_100a code for tail %r (supported: Linear, Constant)" % type(tail).__name__)
