"""Field / evolution / coupling API of the reference, routed to the device stepping core.

Mirrors, for the fixed-step hot path only:
  * gds/fds.py        -- fds.set_evolution / set_initial / set_constraints / step / y / t / dt, coupled_fds, couple()
  * gds/gds.py        -- GraphObservable, gds, node_gds, edge_gds, simplex_gds, face_gds and their operators
  * gds/system.py     -- System.step / solve
Same names, argument meaning and error behaviour; the arithmetic runs in libgds_b200.so (CUDA, sm_100a) and nowhere
else -- there is no CPU fallback.  Operators called with host arrays run eagerly on the device and return fresh
numpy arrays (the `y is not None` path of gds.py:152-345); called while an evolution law is being traced they
return symbolic expressions that are lowered into the captured step (gds_b200/trace.py, engine.py).

`lhs=` laws that are affine in y with a Laplacian-form linear part (the pressure law of examples/fluid.py) are solved
on the device by conjugate gradients.  Out of scope (raise NotImplementedError): general convex `cost=` programmes
(fds.py:97-113), `traj_t=` playback (fds.py:124-129), the adaptive scipy integrators and the dense implicit midpoint
(fds.py:84-95).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib as L
from ._lib import lib, check
from .device import Buffer, BufferView
from .engine import StepEngine
from .graph import LoweredGraph
from .trace import (Binary, Const, DevVec, EdgeAgg, Expr, Gather, Imm, Lowering, MaskZero, PowI, Reduce, Stage, State, Time,
                    TraceError, Unary, as_expr, pass_desc, split_affine, substitute, trace_scope, tracing, unary,
                    _fold_binary, _fold_unary)
from .types import GraphDomain, Integrators, IterationMode, Observable, Steppable
from .utils import dict_fun, fun_ary

_DOMAIN_CODE = {GraphDomain.nodes: L.NODES, GraphDomain.edges: L.EDGES, GraphDomain.faces: L.FACES}


def _identity(x):
    return x


class BoundarySpec:
    """Vectorised boundary condition for graphs too large for a per-point Python callable (extension; the
    reference only takes dicts / callables, fds.py:208-243).  `indices` are rows of the field's domain,
    `values` a constant array, or `fun(t) -> ndarray` for time-varying values."""

    def __init__(self, indices, values=None, fun=None):
        self.indices = np.asarray(indices, dtype=np.intp).ravel()
        assert (values is None) != (fun is None), 'give exactly one of values / fun'
        self.values = None if values is None else np.broadcast_to(np.asarray(values, dtype=float), self.indices.shape).copy()
        self.fun = fun

    def localised(self, part, domain=0) -> 'BoundarySpec':
        """the rows of this rank (owned and ghost) of a condition given on GLOBAL row ids"""
        sel, loc = part.localise(self.indices, domain) if hasattr(part, 'gid_dom') else part.localise(self.indices)
        if self.fun is None:
            return BoundarySpec(loc, values=self.values[sel])
        fun = self.fun
        return BoundarySpec(loc, fun=lambda t: np.asarray(fun(t), dtype=float)[sel])


# ------------------------------------------------------------------------------------------------
# fds (gds/fds.py:14-400)
# ------------------------------------------------------------------------------------------------
class DeviceIntegrator:
    """Solver-object view of a device-stepped field or coupled system: the attributes the reference reads and writes on
    its scipy.integrate.OdeSolver / gds.ode.midpoint.ImplicitMidpoint objects (gds/ode/midpoint.py:10-32;
    fds.py:157,171,196,285-286,362-363,443-444) -- `t`, `y` (the whole state vector: every order block of every coupled
    field, in coupling order; materialised on access, assignable), `t_bound`, `status`, `max_step`, `step()` = ONE
    internal step of min(max_step, t_bound - t) -- over the captured step, for code that drives the solver object
    itself instead of calling `step(dt)`."""

    def __init__(self, owner, members, max_step):
        self._owner, self._members, self._max_step = owner, members, max_step
        self.t_bound = np.inf
        self.status = 'running'

    @property
    def max_step(self):
        return self._max_step()

    @property
    def t(self):
        return self._owner.t

    @property
    def y(self):
        parts = [np.asarray(m._engine.read(m, full=True) if m._engine is not None else m._host_state, dtype=np.float64)
                 for m in self._members]
        return np.concatenate(parts) if len(parts) > 1 else np.array(parts[0], copy=True)

    @y.setter
    def y(self, value):
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float64).ravel())
        assert a.size == sum(m.ndim * m._order() for m in self._members), 'state vector has the wrong length'
        at = 0
        for m in self._members:
            n = m.ndim * m._order()
            if m._engine is not None:
                m._engine.write(m, a[at:at + n])
            else:
                m._host_state = a[at:at + n].copy()
            m._bump()
            at += n

    def step(self):
        if self.status != 'running':
            raise RuntimeError('Attempt to step on a failed or finished solver.')     # scipy's OdeSolver.step
        o = self._owner
        o.step(min(self.max_step, self.t_bound - o.t))
        if self.t_bound - o.t <= 1e-14 * max(1.0, abs(self.t_bound)):      # t + (t_bound - t) may miss t_bound by an ulp
            self.status = 'finished'


class fds(Observable, Steppable):
    def __init__(self, X):
        Observable.__init__(self, X)
        Steppable.__init__(self, IterationMode.none)
        self._set_bcs()
        self.t0 = 0.
        self.y0_fun = lambda _: 0.
        self._engine = None        # StepEngine advancing this field (shared when coupled)
        self._coupled = None       # coupled_fds owning this field
        self._version = 0          # bumped whenever the host-visible state changes
        self._dev_y = self._dev_y_cache = None   # device-resident state of an `lhs=` field

    # ---- evolution law (fds.py:30-135) ---------------------------------------------------------------------
    def set_evolution(self, dydt=None, order=1, max_step=1e-3, integrator=Integrators.rk4, solver_args={},
                      lhs=None, rhs=None, cost=None, refresh_cvx=True, map_fun=None, dt=1.0, traj_t=None, traj_y=None,
                      nil=False):
        given = [dydt is not None, lhs is not None, cost is not None, map_fun is not None,
                 traj_t is not None, bool(nil)]
        assert sum(given) == 1, 'Exactly one evolution law must be specified'       # fds.py:72
        if cost is not None:
            raise NotImplementedError('a general convex cost (cost=, gds/fds.py:97-113) is outside the device stepping path')
        if traj_t is not None:
            raise NotImplementedError('trajectory playback (traj_t=, gds/fds.py:124-129) is outside the device stepping path')
        self._drop_engine()
        if dydt is not None:
            if integrator not in (Integrators.rk4, Integrators.euler):
                # the reference's adaptive scipy solvers / implicit midpoint (fds.py:84-95) are not part of this path
                raise ValueError('Unsupported integrator type')
            self.iter_mode = IterationMode.dydt
            self.dydt_fun = dydt
            self.max_step = max_step
            self._dt = max_step
            self.order = order
            self.solver_args = solver_args
            self.integrator_type = integrator
            self.y0 = np.zeros(self.ndim * order)
        elif lhs is not None:
            # `lhs(t, y) = 0` in the least-squares sense (fds.py:97-113: CVXPY there).  Here: lhs affine in y with a
            # symmetric definite linear part (the pressure-Poisson form of examples/fluid.py:26-31), solved by
            # conjugate gradients on the device (gds_cg_solve).  Parity against CVXPY is unpinned (not installed).
            self.iter_mode = IterationMode.cvx
            self.lhs_fun = lhs
            self.order = 1
            self.y0 = np.zeros(self.ndim)
            self._lhs_cache = {}
            self._dev_y = None
        elif map_fun is not None:
            self.iter_mode = IterationMode.map
            self.map_fun = map_fun            # documented signature (t, y): fds.py:33, README.md:158
            self.order = 1
            self.y0 = np.zeros(self.ndim)
            self._dt = dt
            self._n = self.t0
        else:
            self.iter_mode = IterationMode.nil
            self.order = 1
            self.y0 = np.zeros(self.ndim)
        self._t = self.t0
        # like the reference (midpoint.py:13, fds.py:157-171) the solver state aliases y0 until the first step
        self._host_state = self.y0
        self._bump()

    # ---- initial conditions (fds.py:137-173) -----------------------------------------------------------------
    def set_initial(self, t0=0., y0=lambda _: 0.):
        assert self.iter_mode != IterationMode.none, 'Use set_evolution() before setting initial conditions'
        self._drop_engine()
        self.t0 = t0
        self._t = t0
        if self.iter_mode is IterationMode.map:
            self._n = int(t0)
        if isinstance(y0, np.ndarray):            # the reference recurses here (fds.py:151-152); arrays are accepted
            arr = np.asarray(y0, dtype=float).ravel()
            part = getattr(getattr(self, '_low', None), 'part', None)
            if part is not None and arr.size != self.ndim:
                arr = arr[self.global_ids]            # a global vector: keep this rank's rows
            assert arr.size == self.ndim, 'initial array has the wrong length'
            self.y0_fun = lambda x: arr[self.X[x]]
            keep = np.ones(self.ndim, dtype=bool)
            keep[self.dirichlet_indices] = False
            self._host_state[:self.ndim][keep] = arr[keep]
            if self.y0 is not self._host_state:                 # fds.py:168-173 writes y0 and the solver state
                self.y0[:self.ndim][keep] = arr[keep]
        else:
            self.y0_fun = y0
            constrained = set(self.X_dirichlet)
            st, both = self._host_state, self.y0 is not self._host_state
            for x, i in self.X.items():
                if x not in constrained:
                    v = y0(x)
                    st[i] = v
                    if both:
                        self.y0[i] = v
        self._bump()

    # ---- boundary conditions (fds.py:175-243) ------------------------------------------------------------------
    def set_constraints(self, dirichlet={}, neumann={}, project=_identity):
        assert self.iter_mode != IterationMode.none, 'Use set_evolution() before setting boundary conditions'
        self._drop_engine()
        self._set_bcs(dirichlet, neumann, project)
        # fds.py:193: y0 = project(replace(y0, ...)) -- the replacement is in place, so it also reaches the solver state
        # while that still aliases y0; the projection only does when it works in place
        self.y0[self.dirichlet_indices] = self.dirichlet_values
        self.y0 = np.asarray(project(self.y0), dtype=float)
        if self.iter_mode is IterationMode.dydt:
            # fds.py:196: negative offsets -- for order>1 this addresses the LAST block
            self._host_state[self.dirichlet_indices - self.ndim] = self.dirichlet_values
        elif self.iter_mode in (IterationMode.map, IterationMode.cvx):
            self._host_state = self.y0.copy()                                       # fds.py:197-203
            if self.iter_mode is IterationMode.cvx:
                self._dev_y, self._lhs_cache = None, {}
        else:
            raise Exception('Unsupported iteration mode for set_constraints()')     # fds.py:205
        self._bump()

    def _set_bcs(self, dirichlet={}, neumann={}, project=_identity):
        self.project_fun = project

        def classify(bc):
            if isinstance(bc, BoundarySpec):
                part = getattr(getattr(self, '_low', None), 'part', None)
                if part is not None:
                    bc = bc.localised(part, getattr(self, '_domain_code', 0))   # global row ids: keep this rank's rows
                dyn = bc.fun is not None
                vals = np.asarray(bc.fun(0.), dtype=float).reshape(bc.indices.size) if dyn else bc.values
                fun = (lambda t, x: None) if dyn else (lambda x: None)
                return fun, dyn, None, bc.indices, vals, bc
            if isinstance(bc, dict):
                if not bc:
                    return dict_fun(bc), False, [], np.zeros(0, dtype=np.intp), np.zeros(0), None
                bc = dict_fun(bc)
            dyn = fun_ary(bc) > 1                                                    # fds.py:222-223
            call = (lambda x: bc(0., x)) if dyn else bc
            domain, idx, vals = [], [], []
            for x, i in self.X.items():                                              # fds.py:225-239
                v = call(x)
                if v is not None:
                    domain.append(x); idx.append(i); vals.append(v)
            return bc, dyn, domain, np.asarray(idx, dtype=np.intp), np.asarray(vals, dtype=float).reshape(len(idx)), None

        (self.dirichlet_fun, self.dynamic_dirichlet, self._X_dirichlet, self.dirichlet_indices,
         self.dirichlet_values, self._dirichlet_spec) = classify(dirichlet)
        (self.neumann_fun, self.dynamic_neumann, self._X_neumann, self.neumann_indices,
         self.neumann_values, self._neumann_spec) = classify(neumann)
        both = np.intersect1d(self.dirichlet_indices, self.neumann_indices)
        assert both.size == 0, f'Dirichlet and Neumann conditions overlap on {both}'  # fds.py:242-243

    @property
    def X_dirichlet(self):
        if self._X_dirichlet is None:
            self._X_dirichlet = [self.iX[int(i)] for i in self.dirichlet_indices]
        return self._X_dirichlet

    @property
    def X_neumann(self):
        if self._X_neumann is None:
            self._X_neumann = [self.iX[int(i)] for i in self.neumann_indices]
        return self._X_neumann

    def _dirichlet_at(self, t):
        """values of a time-varying Dirichlet condition at time t (fds.py:350-353)"""
        if self._dirichlet_spec is not None:
            return np.asarray(self._dirichlet_spec.fun(t), dtype=float).reshape(self.dirichlet_indices.size)
        f = self.dirichlet_fun
        return np.fromiter((f(t, x) for x in self.X_dirichlet), dtype=float, count=len(self.X_dirichlet))

    # ---- stepping (fds.py:264-290,332-339) ------------------------------------------------------------------------
    def step(self, dt: float):
        if self._coupled is not None:
            raise Exception('this field is advanced by its coupled system; step that instead')
        if self.iter_mode is IterationMode.none:
            raise Exception('Evolution law not specified')
        elif self.iter_mode is IterationMode.dydt:
            self._ensure_engine().advance(dt)
        elif self.iter_mode is IterationMode.map:
            self._t += dt
            if (self._t - self._n) >= self._dt:
                self._n = self._t
                self._ensure_engine().apply_map(self._t)
        elif self.iter_mode is IterationMode.cvx:
            self._t += dt                     # fds.py:317-328
            self._solve_lhs(self._t)
        elif self.iter_mode is IterationMode.nil:
            self._t += dt
        self._version += 1

    def step_ensemble(self, dt: float, states_in, states_out):
        """Extension (no reference counterpart; it replaces the user's loop `set_initial(y0=m); step(dt); m_new = y` over
        the members of an ensemble, fds.py:137-162,264-290): advance every HOST-resident state in `states_in` (one
        float64 array of ndim * order values per member, pinned memory for full overlap) from t to t + dt and leave the
        results in `states_out`.  Uploads, device steps and read-backs of consecutive members overlap; results are
        bit-identical to the loop.  The field itself ends up holding the last member's state at t + dt."""
        if self._coupled is not None:
            raise Exception('this field is advanced by its coupled system; step that instead')
        if self.iter_mode is not IterationMode.dydt:
            raise NotImplementedError('ensemble stepping covers continuous (dydt=) laws')
        self._ensure_engine().step_ensemble([[a] for a in states_in], [[a] for a in states_out], dt)
        self._version += 1

    def reset(self):
        """back to the initial conditions (fds.py:248-262)"""
        self._drop_engine(pull=False)
        law = {IterationMode.dydt: lambda: dict(dydt=self.dydt_fun, order=self.order, max_step=self.max_step,
                                                integrator=self.integrator_type, solver_args=self.solver_args),
               IterationMode.map: lambda: dict(map_fun=self.map_fun, dt=self._dt),
               IterationMode.cvx: lambda: dict(lhs=self.lhs_fun),
               IterationMode.nil: lambda: dict(nil=True)}[self.iter_mode]()
        d, n, p = self._dirichlet_spec or self.dirichlet_fun, self._neumann_spec or self.neumann_fun, self.project_fun
        t0, y0 = self.t0, self.y0_fun
        self.t0 = 0.
        self.set_evolution(**law)
        self.set_initial(t0=t0, y0=y0)
        if self.iter_mode is not IterationMode.nil:
            self.set_constraints(dirichlet=d, neumann=n, project=p)

    def _order(self):
        return self.order

    def _identity_projection(self):
        return self.project_fun is _identity

    def _state_version(self):
        return (self._version, self._engine.version if self._engine is not None else -1)

    def _bump(self):
        self._version += 1

    def _ensure_engine(self, record=False) -> StepEngine:
        if self._engine is not None and record and not self._engine.record:
            self._drop_engine()                      # rebuild with the snapshot kernel in the captured step
        if self._engine is None:
            scheme = 'map' if self.iter_mode is IterationMode.map else self.integrator_type
            max_step = self.max_step if self.iter_mode is IterationMode.dydt else self._dt
            self._engine = StepEngine([self], scheme, max_step, t0=self._t, record=record)
        return self._engine

    def _drop_engine(self, pull=True):
        """host-side mutation of the set-up: bring the state back and rebuild the captured step lazily"""
        if self._coupled is not None:
            return  # as in the reference, changes after couple() do not reach the coupled integrator (fds.py:171,196)
        if self._engine is not None:
            if pull:
                self._engine.pull_state()
                self._t = self._engine.t
            self._engine = None

    # ---- state (fds.py:373-400) -------------------------------------------------------------------------------------
    @property
    def y(self):
        if tracing():
            return State(self)
        if self.iter_mode is IterationMode.none:
            return np.zeros(self.ndim)
        if self._engine is not None:
            return self._engine.read(self)
        if self.iter_mode is IterationMode.cvx and self._dev_y is not None:
            if self._dev_y_cache is None:
                self._dev_y_cache = self._low.from_dev(self._domain_code, self._dev_y.download())
            return self._dev_y_cache
        return self._host_state[:self.ndim]

    @property
    def t(self):
        if self._coupled is not None:
            return self._coupled.t
        if self.iter_mode is IterationMode.dydt and self._engine is not None:
            return self._engine.t
        return self._t if self.iter_mode is not IterationMode.none else None     # fds.py:385-392: no branch, None

    @property
    def dt(self):
        return self._dt if self.iter_mode in (IterationMode.dydt, IterationMode.map) else 1e-3

    @property
    def integrator(self):
        """the solver object of a `dydt=` field (fds.py:84-95 creates one per field); other modes have none"""
        if self.iter_mode is not IterationMode.dydt:
            raise AttributeError('integrator')
        if self._coupled is not None:
            raise Exception('this field is advanced by its coupled system; that system owns the integrator')
        if self.__dict__.get('_integrator') is None:
            self.__dict__['_integrator'] = DeviceIntegrator(self, [self], lambda: self.max_step)
        return self.__dict__['_integrator']

    def system(self, name: str) -> 'System':
        return System(self, {name: self})


# ------------------------------------------------------------------------------------------------
# fields on graph domains (gds/gds.py)
# ------------------------------------------------------------------------------------------------
class GraphObservable(Observable):
    """index maps of a graph domain (gds.py:16-55); the maps come from the shared lowering of G"""

    def __init__(self, G, Gd: GraphDomain):
        self.G, self.Gd = G, Gd
        self._low = LoweredGraph.of(G, need_faces=(Gd is not GraphDomain.nodes))
        low = self._low
        self.triangles = {}
        self.edge_weights = low.weights
        self._domain_code = _DOMAIN_CODE.get(Gd)
        X = {GraphDomain.nodes: self.nodes, GraphDomain.edges: self.edges, GraphDomain.triangles: self.triangles,
             GraphDomain.faces: self.faces}[Gd]
        Observable.__init__(self, X)

    # index maps live in the shared lowering; `faces` walks the embedding on first use (gds.py:21-36)
    nodes = property(lambda self: self._low.nodes)
    edges = property(lambda self: self._low.edges)
    faces = property(lambda self: self._low.faces)
    outer_face = property(lambda self: self._low.outer_face)
    nodes_i = property(lambda self: {i: v for v, i in self.nodes.items()})
    edges_i = property(lambda self: {i: e for e, i in self.edges.items()})
    faces_i = property(lambda self: {i: f for f, i in self.faces.items()})

    @property
    def edge_orientation(self):   # gds.py:43
        return {**{e: 1 for e in self.edges}, **{(e[1], e[0]): -1 for e in self.edges}}


class gds(fds, GraphObservable):
    def __init__(self, G, Gd: GraphDomain):
        GraphObservable.__init__(self, G, Gd)
        if self._low.part is not None and Gd is not GraphDomain.nodes and self._low.part.hops < 2:
            raise NotImplementedError('edge and face fields need a graph distributed with ghost layers: distribute(G, hops=k), k >= 2')
        fds.__init__(self, self.X)

    # ---- distributed graphs: this rank holds rows [owned | ghosts] -----------------------------------------------
    @property
    def partition(self):
        return self._low.part

    @property
    def global_ids(self):
        """global node id of every row of this field (identity on an undistributed graph)"""
        part = self._low.part
        if part is None:
            return np.arange(self.ndim, dtype=np.int64)
        return part.gid_dom[self._domain_code] if hasattr(part, 'gid_dom') else part.global_ids

    @property
    def owned_mask(self):
        """which rows of this field this rank owns (all of them on an undistributed graph)"""
        part = self._low.part
        if part is None:
            return np.ones(self.ndim, dtype=bool)
        if hasattr(part, 'owned_dom'):
            return part.owned_dom[self._domain_code]
        return np.arange(self.ndim) < part.n_own

    @property
    def y_owned(self):
        """values of the rows this rank owns (all rows on an undistributed graph)"""
        return np.asarray(self.y)[self.owned_mask]

    def global_index(self, x):
        """row of point x in the undistributed field"""
        return int(self.global_ids[self.X[x]])

    # ---- operator plumbing ------------------------------------------------------------------------------------
    def _arg(self, y, domain):
        """operand of an operator: None -> current state (gds.py:153,158,...); host arrays become captured vectors"""
        if y is None:
            y = self.y
        if isinstance(y, gds):
            assert y._low is self._low, 'Incompatible domains'             # gds.py:168,184
            y = y.y
        return y

    def _apply(self, build, *operands, domains):
        """evaluate `build(*exprs)`: symbolically while tracing (or when handed expressions), eagerly otherwise"""
        if tracing() or any(isinstance(o, Expr) for o in operands):
            return build(*[as_expr(o, d, self._low) for o, d in zip(operands, domains)])
        arrs = [np.asarray(o, dtype=float) for o in operands]
        if any(a.ndim == 2 for a in arrs):
            # matrix right-hand side (e.g. laplacian(np.eye(n)), examples/fluid.py:350)
            return _run_eager_cols(self._low, build, arrs, domains)
        for a, d in zip(arrs, domains):
            if a.shape != (self._low.rows(d),):
                raise ValueError(f'operand of shape {a.shape} where a vector of length {self._low.rows(d)} is expected')
        return _run_eager(self._low, build(*[Const(a, d, self._low) for a, d in zip(arrs, domains)]))


def _state_operand(f):
    """a device vector holding the current state of field f, for passes that run outside f's own system"""
    low = f._low
    if f._engine is not None:
        eng = f._engine
        off = eng.offsets[id(f)]
        return BufferView(low.ctx, eng._state_ptr().value + 8 * off, f.ndim)
    if f._dev_y is not None:
        return f._dev_y
    return Buffer(low.ctx, f.ndim, low.to_dev(f._domain_code, f._host_state[:f.ndim]))


def _bind_states(expr, memo):
    """copy of an expression with every State leaf replaced by a device vector of that field's current state"""
    if isinstance(expr, State):
        if id(expr.field) not in memo:
            memo[id(expr.field)] = DevVec(_state_operand(expr.field), expr.domain, expr.graph)
        return memo[id(expr.field)]
    if not expr.args:
        return expr
    import copy
    c = copy.copy(expr)
    c.args = tuple(_bind_states(a, memo) for a in expr.args)
    return c


def _state_fields(*exprs):
    """fields whose accepted state the expressions read"""
    seen, out = set(), []

    def walk(e):
        if e is None or id(e) in seen:
            return
        seen.add(id(e))
        if isinstance(e, State) and not any(e.field is f for f in out):
            out.append(e.field)
        for a in e.args:
            walk(a)
    for e in exprs:
        walk(e)
    return out


def _state_address(f):
    """device address the state of field f is read from (None: still on the host, uploaded per use)"""
    if f._engine is not None:
        return f._engine._state_ptr().value + 8 * f._engine.offsets[id(f)]
    if f._dev_y is not None:
        return id(f._dev_y)
    return None


def _run_eager(low, expr):
    """lower one expression to passes, run them on the device, download the result"""
    ctx = low.ctx
    n = low.rows(expr.domain)
    if n == 0:
        return np.zeros(0)

    def no_leaf(e):
        raise TraceError('field state used outside an evolution law')
    lowering = Lowering(ctx, no_leaf, hoist=False)
    out = Buffer(ctx, n)
    lowering.store_pass(expr, out)
    import ctypes as C
    dev = low.dev
    for p in lowering.passes:
        d, keep = pass_desc(p)
        check(lib.gds_pass_run(ctx.h, dev.h, C.byref(d), 0.0))
    return low.from_dev(expr.domain, out.download())


def _eval_scalar(e, memo):
    """value of a scalar expression whose leaves are constants and device reductions of vector expressions"""
    if isinstance(e, Imm):
        return e.value
    if isinstance(e, Reduce):
        x = e.args[0]
        low = x.graph
        n = low.rows(x.domain)
        if n == 0:
            return 0.0
        if low.part is not None:
            raise TraceError('reductions over a distributed field need a collective')
        bound = _bind_states(x, memo)
        if isinstance(bound, DevVec):                        # a bare state vector: reduce it in place
            return bound.buf.reduce(Reduce.KIND[e.kind], 0, n)
        tmp = Buffer(low.ctx, n)
        descs, keep = _lower_into(low, bound, tmp)
        _run_passes(low, descs)
        return tmp.reduce(Reduce.KIND[e.kind], 0, n)
    if isinstance(e, Binary):
        return _fold_binary(e.op, _eval_scalar(e.args[0], memo), _eval_scalar(e.args[1], memo))
    if isinstance(e, Unary):
        return _fold_unary(e.op, _eval_scalar(e.args[0], memo))
    if isinstance(e, PowI):
        return _eval_scalar(e.args[0], memo) ** e.n
    raise TraceError(f'cannot evaluate a scalar {type(e).__name__}')


def evaluate_view(view, parent):
    """value of a derived observable `parent.project(kind, view)` (gds/types.py:95-105).  The view is first run with symbolic
    field states: if it comes out as an expression over device operators -- `v.curl()`, `(v.y ** 2).sum()`,
    `np.abs(v.curl()).sum()` (gds/examples/lid_driven_cavity.py:95-101) -- its vector part runs on the device against the
    state where it lives (no upload) and a reduction returns 8 bytes; anything else (arbitrary numpy code on `v.y`) runs on
    the host as in the reference."""
    try:
        with trace_scope():
            r = view(parent)
    except Exception:
        r = None
    if isinstance(r, Expr):
        try:
            memo = {}
            if r.domain is None:
                return _eval_scalar(r, memo)
            if r.graph.part is None:
                return _run_eager(r.graph, _bind_states(r, memo))
        except TraceError:
            pass
    return view(parent)


def _run_eager_cols(low, build, arrs, domains):
    """operator on matrix operands, all columns in one device call: one lowering, one upload of every matrix (column
    after column), the pass list enqueued per column with strided operands (gds_passes_run_cols), one download"""
    import ctypes as C
    ctx = low.ctx
    ncol = next(a.shape[1] for a in arrs if a.ndim == 2)
    leaves, strided = [], []
    for a, d in zip(arrs, domains):
        n = low.rows(d)
        if a.ndim == 2:
            if a.shape != (n, ncol):
                raise ValueError(f'operand of shape {a.shape} where a matrix of shape {(n, ncol)} is expected')
            cols = np.ascontiguousarray(low.to_dev(d, a).T)          # [ncol, n]: one contiguous column after the other
            buf = Buffer(ctx, ncol * n, cols)
            strided.append((buf, n))
            leaves.append(DevVec(buf, d, low))
        else:
            if a.shape != (n,):
                raise ValueError(f'operand of shape {a.shape} where a vector of length {n} is expected')
            leaves.append(Const(a, d, low))
    expr = build(*leaves)
    n_out = low.rows(expr.domain)
    if n_out == 0 or ncol == 0:
        return np.zeros((n_out, ncol))
    out = Buffer(ctx, ncol * n_out)
    strided.append((out, n_out))

    def no_leaf(e):
        raise TraceError('field state used outside an evolution law')
    lowering = Lowering(ctx, no_leaf, hoist=False)
    lowering.store_pass(expr, out)
    # intermediate vectors are rewritten for every column: they stay unstrided
    descs = [pass_desc(p) for p in lowering.passes]
    arr = (L.PassDesc * len(descs))(*[d for d, _ in descs])
    bufs = (L.c_vp * len(strided))(*[b.h for b, _ in strided])
    strides = (C.c_int64 * len(strided))(*[int(s) for _, s in strided])
    check(lib.gds_passes_run_cols(ctx.h, low.dev.h, len(descs), arr, 0.0, ncol, len(strided), bufs, strides))
    res = out.download().reshape(ncol, n_out).T
    return np.ascontiguousarray(low.from_dev(expr.domain, res))


def _lower_into(low, expr, out_buf):
    """passes (ctypes descriptors + keep-alives) that evaluate `expr` into the device buffer `out_buf`"""
    def no_leaf(e):
        raise TraceError('field state used outside an evolution law')
    lowering = Lowering(low.ctx, no_leaf, hoist=False)
    lowering.store_pass(expr, out_buf)
    return [pass_desc(p) for p in lowering.passes], lowering


def _run_passes(low, descs):
    import ctypes as C
    dev = low.dev
    for d, keep in descs:
        check(lib.gds_pass_run(low.ctx.h, dev.h, C.byref(d), 0.0))


def _use_multigrid(low, dom, n):
    """node-domain solves on one device are preconditioned with a smoothed-aggregation V-cycle (gds_b200/multigrid.py) from
    4096 rows up; GDS_B200_MG=1 forces it (tests on small graphs), GDS_B200_MG=0 keeps plain conjugate gradients"""
    mode = os.environ.get('GDS_B200_MG', 'auto')
    if dom != L.NODES or low.part is not None or mode == '0':
        return False
    return mode == '1' or n >= 4096


def _cg(low, d_op, x, b, r, p, ap, n, rtol, max_iter, mg=None, z=None):
    """gds_cg_solve, or gds_pcg_solve when a hierarchy is given -> (iterations, relative residual)"""
    import ctypes as C
    import time
    iters, relres = C.c_int32(), C.c_double()
    t0 = time.perf_counter()
    if mg is None:
        check(lib.gds_cg_solve(low.ctx.h, low.dev.h, C.byref(d_op), x.h, b.h, r.h, p.h, ap.h, n, float(rtol), int(max_iter), 25,
                               C.byref(iters), C.byref(relres)))
    else:
        check(lib.gds_pcg_solve(low.ctx.h, low.dev.h, C.byref(d_op), mg.h, x.h, b.h, r.h, p.h, ap.h, z.h, n, float(rtol),
                                int(max_iter), 4, C.byref(iters), C.byref(relres)))
    _cg.last_seconds = time.perf_counter() - t0       # the call returns after its last residual read-back
    return iters.value, relres.value


def _solve_lhs(self, t):
    """`lhs(t, y) = 0` for an lhs affine in y (module docstring of set_evolution): trace once, split into constant and
    linear part, then per call  b = -Z_D(const + lin(g)),  lin(p) = b by conjugate gradients,  y = p + g  (g = Dirichlet
    values, 0 elsewhere).  Everything stays on the device; the result is this field's state."""
    import ctypes as C
    low, ctx, n, dom = self._low, self._low.ctx, self.ndim, self._domain_code
    if self.dynamic_dirichlet:
        self.dirichlet_values = self._dirichlet_at(t)                  # fds.py:350-355
    if 'split' not in self._lhs_cache:
        with trace_scope():
            root = as_expr(self.lhs_fun(Time(), Stage(self)), dom, low)
        const, lin = split_affine(root, self)
        if lin is None:
            raise TraceError('lhs= does not depend on y')
        self._lhs_cache['split'] = (const, lin)
    const, lin = self._lhs_cache['split']
    D = self.dirichlet_indices
    work = self._lhs_cache.setdefault('work', tuple(Buffer(ctx, n) for _ in range(6)))
    x, b, r, p, ap, out = work
    if self._dev_y is None:
        self._dev_y = Buffer(ctx, n)
    # The lowered pass lists are kept between solves.  They bind the other fields' state where it lives, and an engine's
    # accepted state alternates between two buffers, so a plan is keyed by the addresses it was bound to; a field whose
    # state is still on the host (re-uploaded per solve) or time-varying boundary values rule the cache out.
    fields = self._lhs_cache.setdefault('fields', _state_fields(const, lin))
    key = tuple(_state_address(f) for f in fields) + (id(self._dev_y),)
    plans = self._lhs_cache.setdefault('plans', {})
    plan = plans.get(key) if (not self.dynamic_dirichlet and None not in key) else None
    if plan is None:
        g = np.zeros(n)
        g[D] = self.dirichlet_values
        gvec = Const(g, dom, low)
        memo = {}
        rhs = substitute(lin, self, gvec) if const is None else const + substitute(lin, self, gvec)
        rhs = _bind_states(MaskZero(unary('NEG', rhs), D), memo)
        plan = {'rhs': _lower_into(low, rhs, b)}
        op = _bind_states(MaskZero(substitute(lin, self, DevVec(p, dom, low)), D), memo)
        plan['op'] = _lower_into(low, op, ap)
        if len(plan['op'][0]) != 1:
            raise NotImplementedError('lhs=: the part linear in y must be a one-hop operator (a Laplacian form)')
        opn = _bind_states(MaskZero(unary('NEG', substitute(lin, self, DevVec(p, dom, low))), D), memo)
        plan['op_neg'] = _lower_into(low, opn, ap)
        plan['neg_b'] = _lower_into(low, unary('NEG', DevVec(b, dom, low)), r)
        plan['copy_b'] = _lower_into(low, DevVec(r, dom, low) + 0.0, b)
        plan['out'] = _lower_into(low, DevVec(x, dom, low) + gvec, self._dev_y)
        plan['negative'] = False
        if not self.dynamic_dirichlet and None not in key:
            if len(plans) >= 4:
                plans.clear()
            plans[key] = plan
    _run_passes_at(low, plan['rhs'][0], t)
    max_iter = max(1000, 20 * int(np.sqrt(n)) + 200)
    mg = zbuf = None
    if _use_multigrid(low, dom, n):
        # V-cycle of the Dirichlet-masked graph Laplacian as preconditioner: built once per (graph, Dirichlet set); conjugate
        # gradients do not see a positive scaling of the preconditioner, so it serves c * Z_D(L0 .) for every c
        from .multigrid import Multigrid
        mg = Multigrid.for_graph(low, D)
        if 'z' not in self._lhs_cache:
            self._lhs_cache['z'] = Buffer(ctx, n)
        zbuf = self._lhs_cache['z']
    for attempt in range(2):
        d_op = plan['op_neg' if plan['negative'] else 'op'][0]
        if plan['negative']:
            _run_passes_at(low, plan['neg_b'][0], t)
            _run_passes_at(low, plan['copy_b'][0], t)
        iters, relres = _cg(low, d_op[0][0], x, b, r, p, ap, n, 1e-13, max_iter, mg, zbuf)
        if attempt == 0 and not plan['negative'] and iters <= 25 and relres > 0.5:
            # p.Ap was not positive: the linear part is negative definite -- solve (-lin) p = -b instead
            plan['negative'] = True
            continue
        break
    _run_passes_at(low, plan['out'][0], t)
    self._dev_y_cache = None
    self.last_solve = {'iterations': iters, 'relative_residual': relres, 'preconditioner': 'multigrid' if mg else None,
                       'solve_seconds': _cg.last_seconds}
    self._version += 1
    if not self._identity_projection():
        y = np.asarray(self.project_fun(np.array(self.y, copy=True)), dtype=float)   # fds.py:364-366
        self._dev_y.upload(low.to_dev(dom, y))
        self._dev_y_cache = None


fds._solve_lhs = _solve_lhs


def _run_passes_at(low, descs, t):
    import ctypes as C
    dev = low.dev
    for d, keep in descs:
        check(lib.gds_pass_run(low.ctx.h, dev.h, C.byref(d), float(t)))


class node_gds(gds):
    """dynamical system on the nodes of a graph (gds.py:133-189)"""

    def __init__(self, G):
        gds.__init__(self, G, GraphDomain.nodes)
        self.neumann_correction = np.zeros(self.ndim)                               # gds.py:142

    def set_constraints(self, *args, **kwargs):
        fds.set_constraints(self, *args, **kwargs)
        # gds.py:116-120: Dirichlet rows of L0 are zeroed (a row mask here); Neumann values become a constant
        # source that is written once -- time-varying Neumann stays frozen at t=0 like the reference (Appendix B.5)
        self.neumann_correction[self.neumann_indices] = self.neumann_values

    def _lap(self, y):
        out = MaskZero(Gather('N_LAP', self._low, y), self.dirichlet_indices)
        if np.any(self.neumann_correction != 0.):
            out = out + Const(self.neumann_correction, L.NODES, self._low)
        return out

    def partial(self, e, y=None):   # gds.py:149-150
        u, v = e
        w = np.sqrt(self.edge_weights[self.edges[e]])
        if y is None:
            y = self.y
        return w * (y[self.X[v]] - y[self.X[u]])

    def grad(self, y=None):         # gds.py:152-154
        return self._apply(lambda a: Gather('E_GRADT', self._low, a), self._arg(y, L.NODES), domains=(L.NODES,))

    def laplacian(self, y=None):    # gds.py:156-159
        return self._apply(self._lap, self._arg(y, L.NODES), domains=(L.NODES,))

    def bilaplacian(self, y=None):  # gds.py:161-164
        return self._apply(lambda a: MaskZero(Gather('N_LAP', self._low, self._lap(a)), self.dirichlet_indices),
                           self._arg(y, L.NODES), domains=(L.NODES,))

    def advect(self, v_field, y=None):      # gds.py:166-177
        return self._apply(lambda v, a: Gather('N_ADVECT', self._low, v, a), self._arg(v_field, L.EDGES),
                           self._arg(y, L.NODES), domains=(L.EDGES, L.NODES))

    def lie_advect(self, v_field, y=None):  # gds.py:179-189
        return self._apply(lambda v, a: Gather('N_LIE', self._low, v, a), self._arg(v_field, L.EDGES),
                           self._arg(y, L.NODES), domains=(L.EDGES, L.NODES))


class edge_gds(gds):
    """dynamical system on the edges of a graph (gds.py:192-443)"""

    def __init__(self, G, v_free=None):
        gds.__init__(self, G, GraphDomain.edges)
        self.neumann_correction = np.zeros(self.ndim)

    def __call__(self, x):          # gds.py:273-274: oriented read, -1 when x is the reverse of the stored edge
        i = self.X[x]
        stored = self.nodes[x[0]] == self._low.edge_u[i] and self.nodes[x[1]] == self._low.edge_v[i]
        return (1 if stored else -1) * self.y[i]

    def _mask(self, expr, *row_sets):
        rows = [np.asarray(r, dtype=np.int64).ravel() for r in row_sets if r is not None]
        rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
        return MaskZero(expr, rows) if rows.size else expr

    def _dd(self, y, normal=None):
        g = self._low
        return self._mask(unary('NEG', Gather('E_GRADT', g, Gather('N_BDOT', g, y))), self.dirichlet_indices, normal)

    def _d_d(self, y):
        g = self._low
        return self._mask(unary('NEG', Gather('E_CURLT', g, Gather('F_CURL', g, y))), self.dirichlet_indices)

    def _lap(self, y, free=None, normal=None):
        return self._mask(self._dd(y, normal) + self._d_d(y), free)

    def div(self, y=None):          # gds.py:278-280
        return self._apply(lambda a: unary('NEG', Gather('N_BDOT', self._low, a)), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def influx(self, y=None):       # gds.py:282-285
        return self._apply(lambda a: Gather('N_INFLUX', self._low, a), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def outflux(self, y=None):      # gds.py:287-290
        return self._apply(lambda a: Gather('N_OUTFLUX', self._low, a), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def curl(self, y=None):         # gds.py:292-294
        if self._low.F == 0 and not tracing():
            return np.zeros(1)      # the reference keeps a 1 x E zero matrix when there are no faces (gds.py:239)
        return self._apply(lambda a: Gather('F_CURL', self._low, a), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def dd_(self, y=None, normal=None):   # gds.py:296-301
        return self._apply(lambda a: self._dd(a, normal), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def d_d(self, y=None):          # gds.py:303-307
        return self._apply(self._d_d, self._arg(y, L.EDGES), domains=(L.EDGES,))

    def laplacian(self, y=None, free=None, normal=None):   # gds.py:309-317
        return self._apply(lambda a: self._lap(a, free, normal), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def bilaplacian(self, y=None):  # gds.py:319-324
        return self._apply(lambda a: self._lap(self._lap(a)), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def advect(self, y=None, stiff=False, weighted=True):  # gds.py:326-345 (stiff is unused there as well)
        if not weighted:
            raise NotImplementedError('edge_gds.advect(weighted=False) is not part of the device path')
        g = self._low
        return self._apply(lambda a: Gather('E_ADVFIN', g, a, EdgeAgg(g, a)), self._arg(y, L.EDGES), domains=(L.EDGES,))

    def advect_jvp(self, v, y=None):
        """J(y) v, the Jacobian of advect at y applied to the edge vector v (a matrix v: column by column), with the
        flow directions sign(y) held fixed -- the derivative autograd takes through sign() and relu() in advect_torch
        (gds.py:372-395).  Two gather passes like advect itself, no E x E matrix: from advect(y)_e = -(y_e S1_e + s_e S2_e),
            (J v)_e = -v_e S1_e(y) - (y_e S1'_e + s_e S2'_e),
        where S1', S2' are built from the tangent aggregates sum in*w*s_f*v_f and sum in*w*2*y_f*v_f."""
        g = self._low

        def build(a, d):
            return Gather('E_ADVFIN', g, a, EdgeAgg(g, a, d)) - d * Gather('E_ADVS1', g, a, EdgeAgg(g, a))
        return self._apply(build, self._arg(y, L.EDGES), v, domains=(L.EDGES, L.EDGES))

    def advect_jac(self, y=None):
        """Jacobian of the advective derivative as a dense E x E torch tensor (gds.py:391-395, where it is autograd
        through dense operators); assembled from advect_jvp on the identity, so meant for the small graphs the
        reference's Lyapunov analysis uses (examples/turbulence.py:252-268) -- at scale use advect_jvp directly."""
        import torch
        if isinstance(y, torch.Tensor):
            y = y.detach().cpu().numpy()
        y = np.asarray(self._arg(y, L.EDGES), dtype=float)
        return torch.from_numpy(np.ascontiguousarray(self.advect_jvp(np.eye(self.ndim), y)))


    def leray_project(self, y=None, rtol=1e-13, max_iter=None, return_info=False):
        """Projection onto the divergence-free edge fields, y - B1^T pinv(B1 B1^T) B1 y (gds.py:249,414-427).  The
        reference forms the dense |E| x |E| projector with a dense pinv; here B1 B1^T phi = B1 y is solved by conjugate
        gradients on the device (the system is singular -- constants per component -- but consistent, and CG started
        at 0 converges to the pinv solution), then y - B1^T phi.  The reference's branch for fields with Dirichlet
        conditions stacks a non-square system into np.linalg.solve and cannot run; it is not reproduced."""
        import ctypes as C
        if tracing():
            raise TraceError('leray_project solves a linear system; apply it between steps, not inside an evolution law')
        if self.dirichlet_indices.size:
            raise NotImplementedError('leray_project with Dirichlet conditions (gds.py:420-427 is not runnable in the reference)')
        y = np.asarray(self.y if y is None else y, dtype=float)
        if y.shape != (self.ndim,):
            raise ValueError(f'operand of shape {y.shape} where a vector of length {self.ndim} is expected')
        low, ctx = self._low, self._low.ctx
        V, E = low.V, low.E
        if E == 0:
            return y.copy()
        yb = Buffer(ctx, E, y)
        b, x, r, p, ap = (Buffer(ctx, V) for _ in range(5))
        out = Buffer(ctx, E)
        d_rhs, k1 = _lower_into(low, Gather('N_BDOT', low, DevVec(yb, L.EDGES, low)), b)            # B1 y
        _run_passes(low, d_rhs)
        # a right-hand side at rounding level (y already divergence-free) has no meaningful solution: the noise is not in
        # the range of the singular operator
        bh, ynorm = b.download(), float(np.linalg.norm(y))
        if float(np.linalg.norm(bh)) <= 4e-13 * ynorm:
            self.last_solve = {'iterations': 0, 'relative_residual': 0.0}
            return (y.copy(), self.last_solve) if return_info else y.copy()
        # B1 y sums to zero over every connected component; rounding leaves a residue outside the range of B1 B1^T that
        # conjugate gradients cannot reduce -- remove it
        ncomp, lab = low.component_labels()
        bh -= (np.bincount(lab, weights=bh, minlength=ncomp) / np.bincount(lab, minlength=ncomp))[lab]
        b.upload(bh)
        d_op, k2 = _lower_into(low, unary('NEG', Gather('N_LAP', low, DevVec(p, L.NODES, low))), ap)  # B1 B1^T p = -L0 p
        assert len(d_op) == 1
        max_iter = int(max_iter) if max_iter is not None else max(1000, 20 * int(np.sqrt(V)) + 200)
        mg = zbuf = None
        if _use_multigrid(low, L.NODES, V):
            from .multigrid import Multigrid
            mg, zbuf = Multigrid.for_graph(low), Buffer(ctx, V)
        iters, relres = _cg(low, d_op[0][0], x, b, r, p, ap, V, rtol, max_iter, mg, zbuf)
        d_out, k3 = _lower_into(low, DevVec(yb, L.EDGES, low) - Gather('E_GRADT', low, DevVec(x, L.NODES, low)), out)
        _run_passes(low, d_out)
        res = out.download()
        self.last_solve = {'iterations': iters, 'relative_residual': relres, 'preconditioner': 'multigrid' if mg else None,
                       'solve_seconds': _cg.last_seconds}
        return (res, self.last_solve) if return_info else res


class simplex_gds(gds):
    """dynamical system on k-simplices of a graph -- an empty class in the reference as well (gds.py:446-448)"""
    pass


class face_gds(gds):
    """dynamical system on the faces of a graph (gds.py:450-479)"""

    def __init__(self, G):
        gds.__init__(self, G, GraphDomain.faces)

    def laplacian(self, y=None):    # gds.py:473-479: -B2 B2^T y
        g = self._low
        return self._apply(lambda a: unary('NEG', Gather('F_CURL', g, Gather('E_CURLT', g, a))), self._arg(y, L.FACES),
                           domains=(L.FACES,))


# ------------------------------------------------------------------------------------------------
# coupling (gds/fds.py:408-546) and System (gds/system.py:16-57)
# ------------------------------------------------------------------------------------------------
class coupled_fds(Steppable):
    """several fields advanced on ONE concatenated state by one integrator (fds.py:408-507); every field's stage
    of every RK stage runs in the same captured step"""

    def __init__(self, *systems):
        assert len(systems) >= 1, 'Pass one or more systems to couple'
        assert all(s.t == 0. for s in systems), 'All systems must be at zero-time initial conditions.'
        assert all(s.iter_mode != IterationMode.none for s in systems), 'All systems must have evolution laws.'
        Steppable.__init__(self, IterationMode.dydt)
        self.t0 = 0.
        self.cont = [s for s in systems if s.iter_mode is IterationMode.dydt]
        self.maps = [s for s in systems if s.iter_mode is IterationMode.map]
        self.nils = [s for s in systems if s.iter_mode is IterationMode.nil]
        self.cvxs = [s for s in systems if s.iter_mode is IterationMode.cvx]
        if self.cvxs and not self.cont:
            raise NotImplementedError('lhs= fields are coupled to at least one continuous field')
        self.has_integrator = len(self.cont) > 0
        self._engine = None
        self._t = 0.
        self.discrete_t = 0.            # common clock of the discrete systems (fds.py:425)
        members = self.cont or self.maps
        if self.has_integrator:
            kind = self.cont[0].integrator_type
            assert all(s.integrator_type is kind for s in self.cont), 'All continuous dynamics must use the same integrator.'
            self.dydt_max_step = min(s.max_step for s in self.cont)                  # fds.py:437
            self._scheme = kind
        else:
            self._scheme = 'map'
            self.dydt_max_step = min((s._dt for s in self.maps), default=1.0)
        for s in members + (self.maps if self.cont else []):
            s._drop_engine()
            s._host_state = np.array(s.y0, dtype=float, copy=True)                  # fds.py:426-439: state taken at couple() time
            s._couple_y0 = s._host_state.copy()                                      # what reset() goes back to (fds.py:509-527)
            s._coupled = self
        self._members = members
        for s in self.cvxs:
            # the state of an `lhs=` field lives on the device from the start: the continuous laws read it in place
            s._coupled = self
            s._dev_y = Buffer(s._low.ctx, s.ndim, s._low.to_dev(s._domain_code, s._host_state[:s.ndim]))
            s._dev_y_cache = None

    @property
    def integrator(self):
        """the one solver object of the coupled continuous fields (fds.py:441-447)"""
        if not self.has_integrator:
            raise AttributeError('integrator')
        if self.__dict__.get('_integrator') is None:
            self.__dict__['_integrator'] = DeviceIntegrator(self, self.cont, lambda: self.dydt_max_step)
        return self.__dict__['_integrator']

    def _ensure_engine(self, record=False):
        if self._engine is not None and record and not self._engine.record:
            self._engine.pull_state()
            self._t = self._engine.t
            self._engine = None                      # rebuild with the snapshot kernel in the captured step
        if self._engine is None and self._members:
            # recurrence maps may change between two stages of a step (fds.py:505-507): such a step cannot be collapsed
            self._engine = StepEngine(self._members, self._scheme, self.dydt_max_step, t0=self._t, record=record,
                                      allow_collapse=not (self.cont and self.maps))
            for s in self._members:
                s._engine = self._engine
        return self._engine

    # ---- recurrence maps coupled with continuous fields (fds.py:491-507) ------------------------------------------
    def _map_clock(self, dt, commit=True):
        """advance the maps' clocks by dt (fds.py:333-335, step_map); returns the maps that are due.  commit=False: dry run"""
        due = []
        for m in self.maps:
            t = m._t + dt
            if (t - m._n) >= m._dt:
                due.append(m)
                if commit:
                    m._n = t
            if commit:
                m._t = t
        return due

    def _step_discrete(self, dt, t_now):
        """fds.py:491-503: every discrete system advances its clock by dt; a map whose clock has run its own `dt` since
        the last application is applied, at the coupled system's time (`sys.t` is re-bound to it, fds.py:463-467).  All
        maps read the state the others had BEFORE this call (the reference copies the new states after the loop)."""
        due = self._map_clock(dt)
        if due:
            engines = [m._ensure_engine() for m in due]
            for e in engines:
                e._refresh_externals()
            for m, e in zip(due, engines):
                e.apply_map(t_now, refresh=False)
                m._version += 1
        self.discrete_t += dt

    def _step_interleaved(self, dt):
        """fds.py:477-489 with the interleave of fds.py:505-507: the shared integrator calls coupled_fds.dydt at every
        stage, which first steps the discrete systems to the stage time.  Steps during which no map is due replay the
        captured step; a step in which one is due is enqueued stage by stage with the map applied in between."""
        from .engine import step_table
        eng = self._ensure_engine()
        rk4 = self._scheme is Integrators.rk4
        ts, hs, t_new = step_table(eng.t, eng.t + dt, eng.max_step)
        for i in range(ts.size):
            t, h = float(ts[i]), float(hs[i])
            t_next = float(ts[i + 1]) if i + 1 < ts.size else t_new
            for c in self.cvxs:
                c._solve_lhs(t)
            stage_t = (t, t + 0.5 * h, t + 0.5 * h, t + h) if rk4 else (t,)      # the solver's RHS evaluation times
            clock, quiet = self.discrete_t, True
            saved = [(m._t, m._n) for m in self.maps]
            for st_t in stage_t:                      # dry run of the clocks: is any map due inside this step?
                quiet = quiet and not self._map_clock(st_t - clock)
                clock += st_t - clock
            for m, (a, b) in zip(self.maps, saved):
                m._t, m._n = a, b
            if quiet:
                for st_t in stage_t:
                    self._step_discrete(st_t - self.discrete_t, t)
                eng._refresh_externals()
                eng.run((np.asarray([t]), np.asarray([h]), np.asarray([t_next]), t_next, eng._bc_table(np.asarray([t_next]))))
            else:
                def before(stage):
                    self._step_discrete(stage_t[stage] - self.discrete_t, t)
                    eng._refresh_externals()
                eng.run_step_staged(t, h, t_next, before)
            if eng.t > self.discrete_t:               # fds.py:482-483
                self._step_discrete(eng.t - self.discrete_t, eng.t)
            for f in self.cont:
                if not f._identity_projection():       # fds.py:488: arbitrary host code, one round trip
                    full = eng.read(f, full=True)
                    eng.write(f, np.asarray(f.project_fun(full.copy()), dtype=np.float64))
        self._t = eng.t

    def step(self, dt: float):
        if self.has_integrator and self.maps:
            self._step_interleaved(dt)
        elif self.has_integrator and self.cvxs:
            # fds.py:505-507 solves the discrete systems before every RHS evaluation; they read accepted state only, so
            # the solves of one internal step coincide: solve once at its start, then advance that step
            from .engine import step_table
            eng = self._ensure_engine()
            ts, hs, t_new = step_table(eng.t, eng.t + dt, eng.max_step)
            for i in range(ts.size):
                for c in self.cvxs:
                    c._solve_lhs(float(ts[i]))
                eng.advance(float(hs[i]) if i < ts.size - 1 else (t_new - float(ts[i])))
            self._t = eng.t
        elif self.has_integrator:
            self._ensure_engine().advance(dt)
            self._t = self._engine.t
        else:
            self._t += dt
            for s in self.maps:
                s._t = self._t
            if self.maps:
                s0 = self.maps[0]
                if (self._t - s0._n) >= s0._dt:
                    for s in self.maps:
                        s._n = self._t
                    self._ensure_engine().apply_map(self._t)
        for s in self.nils:
            s._t += dt
        for s in self._members:
            s._version += 1

    def step_ensemble(self, dt: float, states_in, states_out):
        """fds.step_ensemble for a coupled system of continuous fields: every member is a list with one array per
        continuous field, in the order of couple()'s arguments"""
        if not self.has_integrator or self.maps or self.cvxs:
            raise NotImplementedError('ensemble stepping covers coupled continuous (dydt=) laws')
        eng = self._ensure_engine()
        eng.step_ensemble([list(m) for m in states_in], [list(m) for m in states_out], dt)
        self._t = eng.t
        for s in self.nils:
            s._t += dt
        for s in self._members:
            s._version += 1

    def reset(self):
        """fds.py:509-527: the shared integrator is rebuilt at t0 on the state taken at couple() time; the discrete systems
        are reset one by one.  (`discrete_t` is left alone there as well.)"""
        if self._engine is not None:
            self._engine = None
            for s in self._members:
                s._engine = None
        self.__dict__.pop('_integrator', None)
        self._t = 0.
        for s in self.cont:
            s._host_state = s._couple_y0.copy()
            s._version += 1
        for s in self.cvxs + self.maps:
            s._coupled = None
            s._engine = None
            s.reset()                                   # fds.py:248-262
            s._coupled = self
            if s.iter_mode is IterationMode.cvx:
                s._dev_y = Buffer(s._low.ctx, s.ndim, s._low.to_dev(s._domain_code, s._host_state[:s.ndim]))
                s._dev_y_cache = None

    def observables(self):
        return self.cont + self.cvxs + self.maps + self.nils

    @property
    def t(self):
        return self._engine.t if (self.has_integrator and self._engine is not None) else self._t


class System:
    def __init__(self, stepper: Steppable, observables):
        self._stepper, self._observables = stepper, observables

    @property
    def stepper(self):
        return self._stepper

    @property
    def observables(self):
        return self._observables

    def step(self, dt: float):
        self._stepper.step(dt)

    def step_ensemble(self, dt: float, states_in, states_out):
        """fds.step_ensemble / coupled_fds.step_ensemble of the system's stepper (extension: an ensemble of HOST-resident
        states advanced by dt with uploads, device steps and read-backs overlapped)"""
        self._stepper.step_ensemble(dt, states_in, states_out)

    def reset(self):
        self._stepper.reset()

    @property
    def t(self):
        return self._stepper.t

    def solve(self, T: float, dt: float, every: int = 1):
        """system.py:38-57; every observable comes back as an ndarray (the reference returns from inside its
        conversion loop, Appendix B.8).  `every` > 1 keeps one snapshot per `every` steps (each snapshot is a
        device-to-host copy)."""
        obs = self._observables
        fast = self._solve_on_device(T, dt, every)
        if fast is not None:
            return fast
        data = {k: [np.array(v.y, copy=True)] for k, v in obs.items()}
        time, t, i = [0.], 0., 0
        while t < T:
            self._stepper.step(dt)
            t += dt
            i += 1
            if i % every == 0 or not (t < T):
                for k, v in obs.items():
                    data[k].append(np.array(v.y, copy=True))
                time.append(t)
        return np.array(time), {k: np.array(v) for k, v in data.items()}


    def solve_to_disk(self, T: float, dt: float, folder: str, parent='runs'):
        """system.py:59-95: step to T and leave one file per observable under parent/folder, each an array
        [n_steps, ndim] of the states AFTER every step (no initial state), plus 't' [n_steps].  The reference writes
        gzip-compressed hickle files; hickle is used when it is importable, otherwise the same arrays go to .npy.
        The trajectory is recorded on the device (solve) and read back once; mesh dumps belong to the renderers."""
        import glob
        import os
        assert os.path.isdir(parent), f'Parent directory "{parent}" does not exist'
        path = parent + '/' + folder
        if not os.path.isdir(path):
            os.mkdir(path)
        else:
            for f in glob.glob(f'{path}/*.hkl') + glob.glob(f'{path}/*.npy'):
                os.remove(f)
        time, data = self.solve(T, dt)
        dump = {'t': time[1:]}
        dump.update({k: v[1:] for k, v in data.items()})
        try:
            import hickle as hkl
        except ImportError:
            hkl = None
        for name, arr in dump.items():
            if hkl is not None:
                hkl.dump(np.asarray(arr), f'{path}/{name}.hkl', mode='w', compression='gzip')
            else:
                np.save(f'{path}/{name}.npy', np.asarray(arr))
        return path

    def _solve_on_device(self, T, dt, every):
        """snapshots taken on the device inside the captured step, one read-back at the end; None when the stepper
        needs the host every step (recurrence maps, user projections, derived observables, distributed graphs)"""
        st, obs = self._stepper, self._observables
        members = [st] if isinstance(st, fds) else list(getattr(st, '_members', []))
        if not members or not hasattr(st, '_ensure_engine'):
            return None
        if isinstance(st, coupled_fds) and (st.cvxs or st.nils or st.maps):
            return None      # solved / advanced / applied between the captured steps: keep the per-step loop
        if any(m.iter_mode is not IterationMode.dydt or not m._identity_projection() or m._low.part is not None
               for m in members):
            return None
        if any(not any(v is m for m in members) for v in obs.values()):
            return None
        time, t, calls = [0.], 0., 0
        while t < T:                       # the reference's loop (system.py:45-50)
            t += dt
            calls += 1
            time.append(t)
        rec = [i for i in range(calls) if (i + 1) % every == 0 or i == calls - 1]
        data = {k: [np.array(v.y, copy=True)] for k, v in obs.items()}
        eng = st._ensure_engine(record=True)
        traj = eng.solve_record(calls, dt, rec)
        if isinstance(st, coupled_fds):
            st._t = eng.t
        for m in members:
            m._version += 1
        for k, v in obs.items():
            off, n = eng.offsets[id(v)], v.ndim
            block = traj[:, off:off + n]
            if v._domain_code == L.NODES and v._low.p2i is not None:
                block = block[:, v._low.p2i]
            data[k] = np.concatenate([data[k][0][None, :], block], axis=0)
        return np.array([time[0]] + [time[i + 1] for i in rec]), data


def couple(observables) -> System:   # fds.py:542-546
    steppables = [o for o in observables.values() if isinstance(o, Steppable)]
    return System(coupled_fds(*steppables), observables)
