"""Tracing of evolution laws and lowering to device passes.

The reference evaluates the user's `dydt(t, y)` / `map_fun(t, y)` lambda in Python at every RHS call
(gds/fds.py:292-305,337), each operator being a numpy matvec (gds/gds.py:152-345).  Here the lambda is run ONCE
with symbolic arguments; the resulting expression DAG is lowered to a short list of *passes* -- one kernel launch
each -- whose row programs (stack code, see include/gds_b200.h) fuse every pointwise operation, the boundary
masking and the Runge-Kutta stage combine into the gather that produces the value.

Leaves distinguish the lambda's `y` argument (the STAGE state) from `obs.y` (the ACCEPTED state): in the reference
`obs.y` is `integrator.y[:ndim]`, which only changes at the end of a step.
"""
from __future__ import annotations

import contextlib
import math
import numbers

import numpy as np

from . import _lib as L
from ._lib import OP

_DOMAIN_NAME = {L.NODES: 'nodes', L.EDGES: 'edges', L.FACES: 'faces', None: 'scalar'}


class TraceError(NotImplementedError):
    pass


# ------------------------------------------------------------------------------------------------
# expression nodes
# ------------------------------------------------------------------------------------------------
class Expr:
    """Symbolic float64 vector over the nodes / edges / faces of a lowered graph, or a scalar (domain None)."""
    __array_priority__ = 1000.0
    domain = None   # L.NODES / L.EDGES / L.FACES / None (scalar)
    graph = None    # LoweredGraph of vector expressions
    args = ()

    # ---- arithmetic ------------------------------------------------------------------------------
    def __add__(self, o): return binary('ADD', self, o)
    def __radd__(self, o): return binary('ADD', o, self)
    def __sub__(self, o): return binary('SUB', self, o)
    def __rsub__(self, o): return binary('SUB', o, self)
    def __mul__(self, o): return binary('MUL', self, o)
    def __rmul__(self, o): return binary('MUL', o, self)
    def __truediv__(self, o): return binary('DIV', self, o)
    def __rtruediv__(self, o): return binary('DIV', o, self)
    def __neg__(self): return unary('NEG', self)
    def __pos__(self): return self
    def __abs__(self): return unary('ABS', self)

    def __pow__(self, o):
        if isinstance(o, numbers.Integral) and -64 <= int(o) <= 64:
            if int(o) == 2:
                return unary('SQUARE', self)
            return PowI(self, int(o))
        return binary('POW', self, o)

    def __rpow__(self, o): return binary('POW', o, self)

    def __bool__(self):
        raise TraceError('the truth value of a traced field expression is not available: data-dependent Python '
                         'control flow cannot be captured into the device step')

    def __len__(self):
        if self.domain is None:
            raise TypeError('scalar expression has no len()')
        return self.graph.rows(self.domain)

    @property
    def shape(self):
        return () if self.domain is None else (len(self),)

    @property
    def size(self):
        return 1 if self.domain is None else len(self)

    ndim = property(lambda self: 0 if self.domain is None else 1)
    dtype = np.dtype('float64')

    def copy(self):
        return self

    # ---- scalar reductions (derived observables: `(v.y ** 2).sum()`, gds/examples/lid_driven_cavity.py:95-101) ----------
    def _reduce(self, kind, axis, out):
        if self.domain is None or axis not in (None, 0) or out is not None:
            raise TraceError('only full reductions of a vector expression can be evaluated on the device')
        return Reduce(kind, self)

    def sum(self, axis=None, dtype=None, out=None, **kw): return self._reduce('sum', axis, out)
    def max(self, axis=None, out=None, **kw): return self._reduce('max', axis, out)
    def min(self, axis=None, out=None, **kw): return self._reduce('min', axis, out)
    def mean(self, axis=None, dtype=None, out=None, **kw): return self._reduce('sum', axis, out) / float(len(self))
    def dot(self, other): return (self * other).sum()

    def __getitem__(self, idx):
        raise TraceError('indexing a traced field expression is not supported in device-captured laws; use the '
                         '`free=` / `normal=` arguments or multiply by a 0/1 vector')

    def __setitem__(self, idx, val):
        """`expr[rows] = 0.` (examples/fluid.py:29): the node turns, in place, into the masked form of itself"""
        if not (isinstance(val, numbers.Real) and float(val) == 0.0) or self.domain is None:
            raise TraceError('only `expr[rows] = 0.` can be captured; multiply by a 0/1 vector for anything else')
        import copy
        inner = copy.copy(self)
        rows = np.arange(len(self))[idx]
        self.__class__ = MaskZero
        self.__dict__.clear()
        MaskZero.__init__(self, inner, rows)

    # ---- numpy ufuncs ------------------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None:
            raise TraceError(f'numpy.{ufunc.__name__}.{method} on traced expressions is not supported')
        name = ufunc.__name__
        if name in _UFUNC_BINARY and len(inputs) == 2:
            return binary(_UFUNC_BINARY[name], inputs[0], inputs[1])
        if name in _UFUNC_UNARY and len(inputs) == 1:
            return unary(_UFUNC_UNARY[name], inputs[0])
        if name == 'power':
            return as_expr(inputs[0]).__pow__(inputs[1]) if isinstance(inputs[0], Expr) else binary('POW', *inputs)
        if name == 'positive':
            return as_expr(inputs[0])
        raise TraceError(f'numpy.{name} is not available inside a device-captured evolution law')


_UFUNC_BINARY = {'add': 'ADD', 'subtract': 'SUB', 'multiply': 'MUL', 'divide': 'DIV', 'true_divide': 'DIV',
                 'maximum': 'MAX', 'minimum': 'MIN', 'arctan2': 'ATAN2', 'fmax': 'MAX', 'fmin': 'MIN',
                 'float_power': 'POW'}
_UFUNC_UNARY = {'negative': 'NEG', 'absolute': 'ABS', 'fabs': 'ABS', 'sign': 'SIGN', 'sqrt': 'SQRT', 'exp': 'EXP',
                'log': 'LOG', 'sin': 'SIN', 'cos': 'COS', 'tan': 'TAN', 'tanh': 'TANH', 'sinh': 'SINH',
                'cosh': 'COSH', 'arctan': 'ATAN', 'square': 'SQUARE', 'reciprocal': 'RECIP'}


class Imm(Expr):
    def __init__(self, v):
        self.value = float(v)


class Time(Expr):
    """the `t` argument of the law: stage time t_n + c_s h"""


class Reduce(Expr):
    """scalar: sum / max / min over the rows of a vector expression, evaluated on the device (gds_buf_reduce)"""
    KIND = {'sum': 0, 'max': 1, 'min': 2}

    def __init__(self, kind, x):
        self.kind, self.args = kind, (x,)


class Stage(Expr):
    """the `y` argument of dydt / map_fun for `field`"""

    def __init__(self, field):
        self.field, self.domain, self.graph = field, field._domain_code, field._low


class State(Expr):
    """`field.y`: accepted state of a field (step start)"""

    def __init__(self, field):
        self.field, self.domain, self.graph = field, field._domain_code, field._low


class Const(Expr):
    """a host vector captured by the law (uploaded once)"""

    def __init__(self, arr, domain, graph):
        self.arr = np.array(arr, dtype=np.float64, copy=True).ravel()
        self.domain, self.graph = domain, graph


class DevVec(Expr):
    """a vector that already lives on the device (work vectors of the solvers)"""

    def __init__(self, buf, domain, graph):
        self.buf, self.domain, self.graph = buf, domain, graph


class Unary(Expr):
    def __init__(self, op, x):
        self.op, self.args, self.domain, self.graph = op, (x,), x.domain, x.graph


class PowI(Expr):
    def __init__(self, x, n):
        self.n, self.args, self.domain, self.graph = n, (x,), x.domain, x.graph


class Binary(Expr):
    def __init__(self, op, a, b):
        self.op, self.args = op, (a, b)
        va = a if a.domain is not None else b
        self.domain, self.graph = va.domain, va.graph


class MaskZero(Expr):
    """x with the rows of `rows` set to zero (Dirichlet rows gds.py:118,299,306; `free`/`normal` gds.py:300,316)"""

    def __init__(self, x, rows):
        self.rows = np.unique(np.asarray(rows, dtype=np.int64).ravel())
        self.args, self.domain, self.graph = (x,), x.domain, x.graph


class Gather(Expr):
    """one-hop incidence operator; `kind` is the opcode name, args are vectors on the source domain(s)"""
    OUT_DOMAIN = {'N_LAP': L.NODES, 'N_BDOT': L.NODES, 'N_INFLUX': L.NODES, 'N_OUTFLUX': L.NODES,
                  'N_ADVECT': L.NODES, 'N_LIE': L.NODES, 'E_GRADT': L.EDGES, 'E_CURLT': L.EDGES,
                  'E_ADVFIN': L.EDGES, 'E_ADVS1': L.EDGES, 'F_CURL': L.FACES}
    IN_DOMAINS = {'N_LAP': (L.NODES,), 'N_BDOT': (L.EDGES,), 'N_INFLUX': (L.EDGES,), 'N_OUTFLUX': (L.EDGES,),
                  'N_ADVECT': (L.EDGES, L.NODES), 'N_LIE': (L.EDGES, L.NODES), 'E_GRADT': (L.NODES,),
                  'E_CURLT': (L.FACES,), 'E_ADVFIN': (L.EDGES, L.NODES), 'E_ADVS1': (L.EDGES, L.NODES),
                  'F_CURL': (L.EDGES,)}

    def __init__(self, kind, graph, *args):
        self.kind, self.graph, self.domain = kind, graph, self.OUT_DOMAIN[kind]
        want = self.IN_DOMAINS[kind]
        assert len(args) == len(want)
        conv = []
        for a, d in zip(args, want):
            a = as_expr(a, d, graph)
            if a.domain is None:
                a = binary('ADD', Const(np.zeros(graph.rows(d)), d, graph), a)  # broadcast a scalar
            if a.domain != d:
                raise ValueError(f'operator {kind} expects a vector over {_DOMAIN_NAME[d]}, got {_DOMAIN_NAME[a.domain]}')
            conv.append(a)
        self.args = tuple(conv)


class EdgeAgg(Expr):
    """the four upwind node aggregates of an edge vector (pass 1 of edge self-advection); 4 values per node.
    With a direction v: their tangent at y along v, the flow directions held fixed (Jacobian-vector product)."""

    def __init__(self, graph, y, v=None):
        self.graph, self.domain = graph, L.NODES
        self.args = (y,) if v is None else (y, v)


def as_expr(x, domain=None, graph=None):
    if isinstance(x, Expr):
        return x
    if isinstance(x, (numbers.Real, np.floating, np.integer)):
        return Imm(x)
    arr = np.asarray(x)
    if arr.dtype == object:
        raise TraceError(f'cannot use {type(x).__name__} inside a device-captured evolution law')
    if arr.ndim == 0:
        return Imm(float(arr))
    if arr.ndim != 1:
        raise TraceError('only 1-D host vectors can be captured by an evolution law')
    if graph is None:
        raise TraceError('cannot infer the domain of a captured host vector')
    if domain is None or graph.rows(domain) != arr.size:
        cands = [d for d in (L.NODES, L.EDGES, L.FACES) if graph.rows(d) == arr.size]
        if domain is not None and domain in cands:
            cands = [domain]
        if len(cands) != 1:
            raise ValueError(f'captured host vector of length {arr.size} does not match a unique domain of the graph')
        domain = cands[0]
    return Const(arr, domain, graph)


def unary(op, x):
    x = as_expr(x)
    if isinstance(x, Imm):
        return Imm(_fold_unary(op, x.value))
    return Unary(op, x)


def binary(op, a, b):
    ea = a if isinstance(a, Expr) else None
    eb = b if isinstance(b, Expr) else None
    ref = ea if (ea is not None and ea.domain is not None) else eb
    dom, gr = (ref.domain, ref.graph) if ref is not None else (None, None)
    a, b = as_expr(a, dom, gr), as_expr(b, dom, gr)
    if a.domain is not None and b.domain is not None:
        if a.domain != b.domain or a.graph is not b.graph:
            raise ValueError(f'cannot combine a vector over {_DOMAIN_NAME[a.domain]} with one over '
                             f'{_DOMAIN_NAME[b.domain]}')
    if isinstance(a, Imm) and isinstance(b, Imm):
        return Imm(_fold_binary(op, a.value, b.value))
    return Binary(op, a, b)


def _fold_unary(op, v):
    f = {'NEG': lambda x: -x, 'ABS': abs, 'SIGN': lambda x: float(np.sign(x)), 'SQRT': math.sqrt, 'EXP': math.exp,
         'LOG': math.log, 'SIN': math.sin, 'COS': math.cos, 'TAN': math.tan, 'TANH': math.tanh, 'SINH': math.sinh,
         'COSH': math.cosh, 'ATAN': math.atan, 'SQUARE': lambda x: x * x, 'RECIP': lambda x: 1.0 / x}[op]
    return f(v)


def _fold_binary(op, a, b):
    f = {'ADD': lambda: a + b, 'SUB': lambda: a - b, 'MUL': lambda: a * b, 'DIV': lambda: a / b,
         'POW': lambda: a ** b, 'MAX': lambda: max(a, b), 'MIN': lambda: min(a, b),
         'ATAN2': lambda: math.atan2(a, b)}[op]
    return f()


# ------------------------------------------------------------------------------------------------
# affine laws: value(y) = const + lin(y)   (the `lhs=` form of set_evolution)
# ------------------------------------------------------------------------------------------------
_LINEAR_GATHERS = ('N_LAP', 'N_BDOT', 'E_GRADT', 'E_CURLT', 'F_CURL')


def depends_on(e, field):
    if isinstance(e, Stage):
        return e.field is field
    return any(depends_on(a, field) for a in e.args)


def substitute(e, field, replacement):
    """copy of e with the stage vector of `field` replaced"""
    if isinstance(e, Stage) and e.field is field:
        return replacement
    if not e.args or not depends_on(e, field):
        return e
    import copy
    c = copy.copy(e)
    c.args = tuple(substitute(a, field, replacement) for a in e.args)
    return c


def split_affine(e, field):
    """(const, lin): e == const + lin with lin linear in the stage vector of `field` (either may be None)"""
    if not depends_on(e, field):
        return e, None
    if isinstance(e, Stage):
        return None, e

    def add(a, b, op='ADD'):
        if a is None:
            return b if op == 'ADD' or b is None else unary('NEG', b)
        return a if b is None else binary(op, a, b)
    if isinstance(e, Binary) and e.op in ('ADD', 'SUB'):
        (ca, la), (cb, lb) = split_affine(e.args[0], field), split_affine(e.args[1], field)
        return add(ca, cb, e.op), add(la, lb, e.op)
    if isinstance(e, Binary) and e.op == 'MUL':
        a, b = e.args
        if depends_on(a, field) and depends_on(b, field):
            raise TraceError('lhs= must be affine in y (a product of two y-dependent factors was found)')
        dep, ind = (a, b) if depends_on(a, field) else (b, a)
        c, l = split_affine(dep, field)
        return (None if c is None else binary('MUL', c, ind)), (None if l is None else binary('MUL', l, ind))
    if isinstance(e, Binary) and e.op == 'DIV' and not depends_on(e.args[1], field):
        c, l = split_affine(e.args[0], field)
        return (None if c is None else binary('DIV', c, e.args[1])), (None if l is None else binary('DIV', l, e.args[1]))
    if isinstance(e, Unary) and e.op == 'NEG':
        c, l = split_affine(e.args[0], field)
        return (None if c is None else unary('NEG', c)), (None if l is None else unary('NEG', l))
    if isinstance(e, MaskZero):
        c, l = split_affine(e.args[0], field)
        return (None if c is None else MaskZero(c, e.rows)), (None if l is None else MaskZero(l, e.rows))
    if isinstance(e, Gather) and e.kind in _LINEAR_GATHERS:
        c, l = split_affine(e.args[0], field)
        return (None if c is None else Gather(e.kind, e.graph, c)), (None if l is None else Gather(e.kind, e.graph, l))
    raise TraceError(f'lhs= must be affine in y ({type(e).__name__} of y is not)')


# ------------------------------------------------------------------------------------------------
# trace context: while active, `field.y` returns a State leaf and operators build expressions
# ------------------------------------------------------------------------------------------------
_ACTIVE = []


def tracing():
    return bool(_ACTIVE)


@contextlib.contextmanager
def trace_scope():
    _ACTIVE.append(True)
    try:
        yield
    finally:
        _ACTIVE.pop()


# ------------------------------------------------------------------------------------------------
# lowering
# ------------------------------------------------------------------------------------------------
class Pass:
    def __init__(self, domain, when, graph):
        self.domain, self.when, self.graph = domain, when, graph
        self.code, self.imm, self.vec, self.mask, self.out = [], [], [], [], []
        self.epilogue, self.state_offset, self.dirichlet = L.EPI_NONE, 0, None
        self.depth = 0
        self.max_depth = 0
        self._vec_index = {}
        self._imm_index = {}
        self._mask_index = {}

    def emit(self, op, a=0, b=0, push=0, pop=0, c=0):
        self.code.append(OP[op] | ((a & 0xff) << 8) | ((b & 0xff) << 16) | ((c & 0xff) << 24))
        self.depth += push - pop
        self.max_depth = max(self.max_depth, self.depth)
        if len(self.code) > L.MAX_CODE:
            raise TraceError('evolution law too long for one row program; split it with intermediate fields')

    def vec_slot(self, key, ref):
        if key not in self._vec_index:
            if len(self.vec) >= L.MAX_VEC:
                raise TraceError('evolution law reads too many distinct vectors in one pass')
            self._vec_index[key] = len(self.vec)
            self.vec.append(ref)
        return self._vec_index[key]

    def imm_slot(self, v):
        key = np.float64(v).tobytes()
        if key not in self._imm_index:
            if len(self.imm) >= L.MAX_IMM:
                raise TraceError('evolution law uses too many distinct constants in one pass')
            self._imm_index[key] = len(self.imm)
            self.imm.append(float(v))
        return self._imm_index[key]

    def mask_slot(self, key, mask):
        if key not in self._mask_index:
            if len(self.mask) >= L.MAX_MASK:
                raise TraceError('evolution law uses too many distinct row masks in one pass')
            self._mask_index[key] = len(self.mask)
            self.mask.append(mask)
        return self._mask_index[key]


class Lowering:
    """Lower expression DAGs to passes.  `resolve_leaf(expr)` maps Stage/State leaves to operand references
    (kind, Buffer|None, offset); the caller owns that policy (system layout)."""

    def __init__(self, ctx, resolve_leaf, hoist=True, single_hop=False):
        from .device import Buffer, Mask
        self.single_hop = single_hop   # partitioned graphs: ghost rows of a gathered vector are not valid gather inputs
        self._Buffer, self._Mask = Buffer, Mask
        self.ctx = ctx
        self.resolve_leaf = resolve_leaf
        self.hoist = hoist
        self.passes = []
        self._temps = {}      # id(expr) -> (Buffer, when)
        self._consts = {}     # id(Const) -> Buffer
        self._masks = {}      # (domain, bytes) -> Mask
        self._variant = {}    # id(expr) -> bool (depends on stage state or time)
        self._need = {}
        self._keep = []       # keep expressions alive so id() keys stay unique
        self.temp_bytes = 0

    # -- analysis ---------------------------------------------------------------------------------
    def variant(self, e):
        k = id(e)
        if k not in self._variant:
            self._keep.append(e)
            if isinstance(e, (Stage, Time)):
                v = True
            else:
                v = any(self.variant(a) for a in e.args)
            self._variant[k] = v
        return self._variant[k]

    def has_gather(self, e):
        if isinstance(e, (Gather, EdgeAgg)):
            return True
        return any(self.has_gather(a) for a in e.args)

    def need(self, e):
        """stack slots needed to evaluate e (Sethi-Ullman)"""
        k = id(e)
        if k not in self._need:
            if isinstance(e, Binary):
                na, nb = self.need(e.args[0]), self.need(e.args[1])
                n = max(na, nb) if na != nb else na + 1
            elif isinstance(e, (Unary, PowI, MaskZero)):
                n = self.need(e.args[0])
            else:
                n = 1
            self._need[k] = n
        return self._need[k]

    # -- operands ----------------------------------------------------------------------------------
    def operand(self, p, e, for_gather=False):
        """vector slot in pass p for a vector-valued expression that must live in memory"""
        if for_gather and self.single_hop and self.has_gather(e):
            raise TraceError('two-hop operators (a gather of a gathered vector: bilaplacian, Hodge Laplacian, ...) are '
                             'not available on a distributed graph yet: only one halo layer is exchanged per stage')
        if isinstance(e, (Stage, State)):
            ref = self.resolve_leaf(e)
            if ref[0] == L.VEC_STAGE and p.when == L.WHEN_STEP:
                raise TraceError('internal: stage operand in a per-step pass')
            return p.vec_slot(('leaf', ref[0], id(ref[1]), ref[2]), ref)
        if isinstance(e, DevVec):
            return p.vec_slot(('buf', id(e.buf), 0), (L.VEC_BUF, e.buf, 0))
        if isinstance(e, Const):
            if id(e) not in self._consts:
                self._keep.append(e)
                self._consts[id(e)] = self._Buffer(self.ctx, e.arr.size, e.graph.to_dev(e.domain, e.arr))
            buf = self._consts[id(e)]
            return p.vec_slot(('buf', id(buf), 0), (L.VEC_BUF, buf, 0))
        buf = self.materialize(e)
        return p.vec_slot(('buf', id(buf), 0), (L.VEC_BUF, buf, 0))

    def materialize(self, e):
        """compute e into a temp vector with its own pass (memoised)"""
        k = id(e)
        if k in self._temps:
            return self._temps[k][0]
        self._keep.append(e)
        when = L.WHEN_STAGE if (self.variant(e) or not self.hoist) else L.WHEN_STEP
        rows = e.graph.rows(e.domain)
        if isinstance(e, EdgeAgg):
            buf = self._Buffer(self.ctx, 4 * rows)
            p = Pass(L.NODES, when, e.graph)
            ys = self.operand(p, e.args[0], for_gather=True)
            p.out.append((buf, 0))
            if len(e.args) == 2:
                p.emit('N_EAGGT', ys, 0, c=self.operand(p, e.args[1], for_gather=True))
            else:
                p.emit('N_EAGG', ys, 0)
        else:
            buf = self._Buffer(self.ctx, rows)
            p = Pass(e.domain, when, e.graph)
            self.emit(p, e)
            p.out.append((buf, 0))
            p.emit('STORE', 0, 0, push=1, pop=1)
        self.temp_bytes += buf.n * 8
        self._check(p)
        self.passes.append(p)
        self._temps[k] = (buf, when)
        return buf

    def _check(self, p):
        if p.max_depth > L.MAX_STACK:
            raise TraceError('evolution law nests too deeply for the row-program stack; split it with intermediate fields')

    # -- code generation ---------------------------------------------------------------------------
    def emit(self, p, e):
        """append code to pass p leaving the value of e (at the pass's row) on the stack"""
        if isinstance(e, Imm):
            p.emit('PUSH_IMM', p.imm_slot(e.value), push=1)
            return
        if isinstance(e, Time):
            p.emit('PUSH_T', push=1)
            return
        if e.domain is not None and e.domain != p.domain:
            raise ValueError(f'a vector over {_DOMAIN_NAME[e.domain]} cannot be used pointwise in a law over '
                             f'{_DOMAIN_NAME[p.domain]}')
        # hoist accepted-state-only sub-expressions that contain a gather out of the per-stage pass
        if (self.hoist and p.when == L.WHEN_STAGE and not isinstance(e, (Stage, State, Const, DevVec))
                and e.domain is not None and not self.variant(e) and self.has_gather(e)):
            p.emit('PUSH_VEC', self.operand(p, e), push=1)
            return
        if id(e) in self._temps and e.domain is not None:
            p.emit('PUSH_VEC', self.operand(p, e), push=1)
            return
        if isinstance(e, (Stage, State, Const, DevVec)):
            p.emit('PUSH_VEC', self.operand(p, e), push=1)
        elif isinstance(e, Unary):
            self.emit(p, e.args[0])
            p.emit(e.op, push=1, pop=1)
        elif isinstance(e, PowI):
            self.emit(p, e.args[0])
            p.emit('POWI', e.n & 0xff, push=1, pop=1)
        elif isinstance(e, Binary):
            a, b = e.args
            if self.need(b) > self.need(a):
                self.emit(p, b)
                self.emit(p, a)
                if e.op not in ('ADD', 'MUL', 'MAX', 'MIN'):
                    p.emit('SWAP', push=2, pop=2)
            else:
                self.emit(p, a)
                self.emit(p, b)
            p.emit(e.op, push=1, pop=2)
        elif isinstance(e, MaskZero):
            self.emit(p, e.args[0])
            if e.rows.size:
                key = (e.domain, id(e.graph), e.rows.tobytes())
                if key not in self._masks:
                    self._masks[key] = self._Mask(self.ctx, e.graph.rows(e.domain), e.graph.rows_to_dev(e.domain, e.rows))
                p.emit('MASK_ZERO', p.mask_slot(key, self._masks[key]), push=1, pop=1)
        elif isinstance(e, Gather) and e.kind == 'N_LAP' and self._chain_of(e.args[0]) is not None:
            # L0 f(v) with f a short pointwise chain over a vector in memory: evaluated on the fly inside the gather
            # (for the row and each neighbour) instead of through an intermediate vector
            leaf, steps = self._chain_of(e.args[0])
            base = len(p.imm)
            if base + 2 * len(steps) > L.MAX_IMM:
                raise TraceError('evolution law uses too many distinct constants in one pass')
            for code, k in steps:
                p.imm.extend([float(code), float(k)])
            p.code.append(OP['N_LAP'] | (self.operand(p, leaf, for_gather=True) << 8) | (base << 16) | (len(steps) << 24))
            p.depth += 1
            p.max_depth = max(p.max_depth, p.depth)
        elif isinstance(e, Gather):
            slots = [self.operand(p, a, for_gather=True) for a in e.args]
            if e.kind in ('E_ADVFIN', 'E_ADVS1'):
                p.emit(e.kind, slots[0], slots[1], push=1)
            elif len(slots) == 2:
                p.emit(e.kind, slots[0], slots[1], push=1)
            else:
                p.emit(e.kind, slots[0], push=1)
        elif isinstance(e, EdgeAgg):
            raise TraceError('internal: edge aggregates are only consumed by E_ADVFIN / E_ADVS1')
        else:
            raise TraceError(f'cannot lower {type(e).__name__}')

    CHAIN = {'SQUARE': 1, 'MUL': 2, 'ADD': 3, 'RSUB': 4, 'NEG': 5, 'ABS': 6, 'CUBE': 7}   # GDS_CHAIN_*

    def _chain_of(self, e):
        """(leaf, [(code, constant), ...]) if e is a chain of at most 4 pointwise steps with constants over ONE vector
        that lives in memory, in evaluation order; None otherwise (or for a bare leaf: nothing to fuse)"""
        steps = []
        while True:
            if isinstance(e, (Stage, State, Const, DevVec)):
                break
            if isinstance(e, Unary) and e.op in ('SQUARE', 'NEG', 'ABS'):
                steps.append((self.CHAIN[e.op], 0.0))
                e = e.args[0]
            elif isinstance(e, PowI) and e.n == 3:
                steps.append((self.CHAIN['CUBE'], 0.0))
                e = e.args[0]
            elif isinstance(e, Binary) and e.op in ('ADD', 'SUB', 'MUL'):
                a, b = e.args
                if isinstance(b, Imm) and not isinstance(a, Imm):
                    # x op k:  x + k, x - k == x + (-k) exactly, x * k
                    code, k = {'ADD': ('ADD', b.value), 'SUB': ('ADD', -b.value), 'MUL': ('MUL', b.value)}[e.op]
                    steps.append((self.CHAIN[code], k))
                    e = a
                elif isinstance(a, Imm) and e.op in ('ADD', 'SUB', 'MUL'):
                    code = {'ADD': 'ADD', 'SUB': 'RSUB', 'MUL': 'MUL'}[e.op]       # k + x, k - x, k * x (commutative: exact)
                    steps.append((self.CHAIN[code], a.value))
                    e = b
                else:
                    return None
            else:
                return None
            if len(steps) > 4:
                return None
        if not steps or id(e) in self._temps:
            return None
        return e, steps[::-1]

    def rhs_pass(self, e, domain, graph, state_offset, dirichlet_mask):
        """final pass of a field: value of e is dy/dt of the rows at state_offset"""
        e = as_expr(e, domain, graph)
        p = Pass(domain, L.WHEN_STAGE, graph)
        self.emit(p, e)
        p.epilogue, p.state_offset, p.dirichlet = L.EPI_RHS, state_offset, dirichlet_mask
        self._check(p)
        self.passes.append(p)
        return p

    def store_pass(self, e, out_buf):
        """eager application: evaluate e into out_buf"""
        p = Pass(e.domain, L.WHEN_STAGE, e.graph)
        self.emit(p, e)
        p.out.append((out_buf, 0))
        p.emit('STORE', 0, 0, push=1, pop=1)
        self._check(p)
        self.passes.append(p)
        return p


def pass_desc(p: Pass):
    """ctypes gds_pass_desc for a Pass; returns (desc, keepalive)."""
    import ctypes as C
    d = L.PassDesc()
    code = (C.c_uint32 * max(1, len(p.code)))(*p.code)
    imm = (C.c_double * max(1, len(p.imm)))(*p.imm)
    vec = (L.VecRef * max(1, len(p.vec)))()
    for i, (kind, buf, off) in enumerate(p.vec):
        vec[i].kind, vec[i].buf, vec[i].offset = kind, (buf.h if buf is not None else None), int(off)
    mask = (L.c_vp * max(1, len(p.mask)))(*[m.h for m in p.mask])
    out = (L.c_vp * max(1, len(p.out)))(*[b.h for b, _ in p.out])
    out_off = (C.c_int64 * max(1, len(p.out)))(*[int(o) for _, o in p.out])
    d.domain, d.when = p.domain, p.when
    d.n_code, d.code = len(p.code), code
    d.n_imm, d.imm = len(p.imm), imm
    d.n_vec, d.vec = len(p.vec), vec
    d.n_mask, d.mask = len(p.mask), mask
    d.n_out, d.out, d.out_offset = len(p.out), out, out_off
    d.epilogue, d.state_offset = p.epilogue, int(p.state_offset)
    d.dirichlet = p.dirichlet.h if p.dirichlet is not None else None
    return d, (code, imm, vec, mask, out, out_off)
