"""Stand-in for the subset of CVXPY that the reference's `lhs=` / `cost=` path uses -- TEST INFRASTRUCTURE ONLY.

CVXPY is not installed in this image, so `set_evolution(lhs=...)` of the UNMODIFIED reference (gds/fds.py:97-113,
196-202, 307-328: the pressure field of gds/examples/fluid.py:14-38) could not run and parity of that mode was unpinned.
This module lets the reference's own code run: `oracle/ref_harness.py` installs it as `cvxpy` before importing the
reference.  It implements exactly what those lines touch --

    cp.Parameter(nonneg=True) / cp.Parameter(n)      .value
    cp.Variable(n)                                   .value, y[idx], A @ y (numpy / scipy.sparse A), + - * / with arrays
    cp.sum_squares(expr)                             .shape == (), .is_dcp()
    cp.Minimize(cost), cp.Problem(objective, constraints)   .objective, .constraints, .solve(**kw), .status
    y[idx] == parameter                              (the Dirichlet constraint of fds.py:199-202)

-- with expressions kept as affine maps `M y + c` of the one variable, and `Problem.solve` minimising `|M y + c|^2`
subject to the equality constraints EXACTLY (constrained entries eliminated, dense least squares on the rest; the
minimum-norm minimiser when it is not unique).  CVXPY would hand the same strictly convex quadratic programme to OSQP /
ECOS and return its minimiser to solver tolerance; what runs unmodified here -- and what the golden vectors made with this
stand-in pin -- is everything the REFERENCE contributes: how `lhs(t, y)` is assembled from its operators, the Dirichlet
constraint, when the solves happen relative to the Runge-Kutta stages of a coupled system (fds.py:491-507), and the state
bookkeeping around them.  Nothing in the product path imports this module.
"""
import numpy as np
import scipy.sparse as sp


def _as_matrix(a):
    return a if sp.issparse(a) else np.asarray(a, dtype=float)


class Affine:
    """M y + c for the problem's single variable y (M: dense (m, n) array, c: (m,) array)"""
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """numpy operators with an ndarray on the left (`a + e`, `a - e`, `a * e`, `a / e`, `A @ e`) and their in-place forms:
        `lhs -= pressure.laplacian(y) / density` (gds/examples/fluid.py:30) has an ndarray on the left; numpy never defers
        an in-place operator, it calls this hook with out=(lhs,) -- the result is the new expression, which Python then
        binds to the name (`lhs` becomes an expression, which is what the example means)."""
        if method != '__call__' or len(inputs) != 2:
            return NotImplemented
        a, b = inputs
        if b is not self or isinstance(a, Affine):
            return NotImplemented                       # expression on the left: the ordinary methods have run already
        if ufunc is np.add:
            return self.__radd__(a)
        if ufunc is np.subtract:
            return self.__rsub__(a)
        if ufunc is np.multiply:
            return self.__rmul__(a)
        if ufunc is np.matmul:
            return self.__rmatmul__(a)
        return NotImplemented

    def __init__(self, var, M, c):
        self.var, self.M, self.c = var, np.asarray(M, dtype=float), np.asarray(c, dtype=float)

    @property
    def shape(self):
        return self.c.shape

    @property
    def size(self):
        return self.c.size

    def _const(self, other):
        """other as a constant vector broadcast to this expression's shape, or None when it is an expression"""
        if isinstance(other, Affine):
            return None
        if isinstance(other, Parameter):
            other = other.value
        return np.broadcast_to(np.asarray(other, dtype=float), self.c.shape)

    def __add__(self, other):
        k = self._const(other)
        if k is None:
            assert other.var is self.var, 'one variable per problem'
            return Affine(self.var, self.M + other.M, self.c + other.c)
        return Affine(self.var, self.M, self.c + k)

    __radd__ = __add__

    def __neg__(self):
        return Affine(self.var, -self.M, -self.c)

    def __sub__(self, other):
        return self + (-other if isinstance(other, Affine) else -np.asarray(other.value if isinstance(other, Parameter) else other, dtype=float))

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if isinstance(other, Parameter):
            other = other.value
        k = np.asarray(other, dtype=float)
        if k.ndim == 0:
            return Affine(self.var, self.M * float(k), self.c * float(k))
        k = np.broadcast_to(k, self.c.shape)            # elementwise scaling of the rows
        return Affine(self.var, self.M * k[:, None], self.c * k)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Parameter):
            other = other.value
        return self * (1.0 / np.asarray(other, dtype=float))

    def __rmatmul__(self, A):
        A = _as_matrix(A)
        return Affine(self.var, np.asarray(A @ self.M), np.asarray(A @ self.c).ravel())

    def __getitem__(self, idx):
        return Affine(self.var, self.M[idx], self.c[idx])

    def __setitem__(self, idx, value):
        """`lhs[rows] = 0.` of gds/examples/fluid.py:29 applied to an expression"""
        if isinstance(value, Affine):
            self.M[idx], self.c[idx] = value.M, value.c
        else:
            self.M[idx] = 0.0
            self.c[idx] = value

    def __eq__(self, other):                            # noqa: A003 -- builds a constraint, as in CVXPY
        return Equality(self, other)

    __hash__ = None


class Variable(Affine):
    def __init__(self, n):
        n = int(n)
        Affine.__init__(self, self, np.eye(n), np.zeros(n))
        self.value = None


class Parameter:
    def __init__(self, shape=(), nonneg=False, **kw):
        self.shape = (int(shape),) if np.isscalar(shape) and not isinstance(shape, tuple) else tuple(shape)
        self.value = None


class Equality:
    def __init__(self, expr, rhs):
        self.expr, self.rhs = expr, rhs

    def rhs_value(self):
        r = self.rhs.value if isinstance(self.rhs, Parameter) else self.rhs
        return np.broadcast_to(np.asarray(r, dtype=float), self.expr.c.shape)


class SumSquares:
    """scalar cost |expr|^2"""
    shape = ()

    def __init__(self, expr):
        self.expr = expr

    def is_dcp(self):
        return True


def sum_squares(expr):
    if isinstance(expr, SumSquares):
        return expr
    if not isinstance(expr, Affine):
        raise TypeError('cvx_shim.sum_squares: the argument does not depend on the variable')
    return SumSquares(expr)


class Minimize:
    def __init__(self, cost):
        if not isinstance(cost, SumSquares):
            raise TypeError('cvx_shim supports least-squares objectives (sum_squares of an affine expression) only')
        self.cost = cost


class Problem:
    def __init__(self, objective, constraints=()):
        self.objective, self.constraints = objective, list(constraints)
        self.status = None
        self.value = None

    def solve(self, **kwargs):
        """argmin |M y + c|^2 subject to the equality constraints, exactly"""
        e = self.objective.cost.expr
        var = e.var
        n = var.c.size
        fixed = np.zeros(n, dtype=bool)
        yfix = np.zeros(n)
        for con in self.constraints:
            assert isinstance(con, Equality) and con.expr.var is var
            Mc, rhs = con.expr.M, con.rhs_value() - con.expr.c
            # the reference only ever constrains entries of the variable: rows of the identity
            rows, cols = np.nonzero(Mc)
            assert rows.size == Mc.shape[0] and np.array_equal(rows, np.arange(Mc.shape[0])) and np.all(Mc[rows, cols] == 1.0), \
                'cvx_shim supports constraints of the form y[indices] == values'
            fixed[cols] = True
            yfix[cols] = rhs
        free = ~fixed
        y = yfix.copy()
        if free.any():
            sol, *_ = np.linalg.lstsq(e.M[:, free], -(e.c + e.M[:, fixed] @ yfix[fixed]), rcond=None)
            y[free] = sol
        var.value = y
        r = e.M @ y + e.c
        self.value = float(r @ r)
        self.status = 'optimal'
        return self.value
