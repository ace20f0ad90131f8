"""CPU oracle for the gds hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy.sparse restatement of the reference's (asrvsn/gds) fixed-step integration path:
operator assembly, operator application, boundary-condition bookkeeping and the stepping loop.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import this module; `gds_b200/` never does (the product path is CUDA-only and fails loudly
without its extension).

Parity status: PINNED.  `tests/golden/make_golden.py` runs the UNMODIFIED reference (imported
from /root/reference through `oracle/ref_harness.py`) on small graphs and stores its outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this restatement against them.  The
reference ships no tests or golden vectors of its own.

The reference keeps its operators as DENSE |V|x|E| / |E|x|E| ndarrays (gds/gds.py:145,265-267);
this restatement keeps exactly the same matrices in scipy CSR form, so the arithmetic is the same
linear algebra with the zeros skipped.  Every function cites the reference lines it follows.

Deliberate positions on reference quirks:
  * fixed-step `RK4`/`Euler` solver objects are NEW (the reference has only adaptive scipy solvers and
    an implicit midpoint); they follow the solver-object protocol of gds/ode/midpoint.py:10-32.
  * map mode follows gds/fds.py:332-339 with the two fixes needed to make it reachable
    (`_y = y0.copy()` for fds.py:122, documented `(t, y)` signature for fds.py:337).
  * `set_initial(y0=ndarray)` is accepted (fds.py:151-152 recurses in the reference).
  * `lhs=` (cvx mode, fds.py:97-113,309-328): the reference hands `sum_squares(lhs(t, y))` to CVXPY; for an `lhs` affine
    in y the restatement takes the least-squares solution with the Dirichlet values imposed (minimum norm over the free
    unknowns).  CVXPY is not installed here.  PARITY PINNED TO THE REFERENCE'S OWN CVX CODE PATH, run unmodified with
    oracle/cvx_shim.py standing in for `cvxpy` (affine expressions of the variable; the programme is minimised exactly by
    dense least squares): tests/golden/lhs_*.npz, <= 1e-12.  What that pins is everything the reference contributes --
    how the objective is assembled from its operators, the Dirichlet constraint, the solves before every RHS evaluation of
    a coupled system, the state bookkeeping; what it cannot pin is CVXPY's own numerical solve (OSQP / ECOS tolerance).
"""
from __future__ import annotations

import inspect
import math
from enum import Enum

import networkx as nx
import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------
# enums (gds/types.py:23-51; rk4/euler are the new fixed-step members)
# --------------------------------------------------------------------------------------------
class GraphDomain(Enum):
    nodes = 0
    edges = 1
    triangles = 2
    faces = 3


class IterationMode(Enum):
    none = 0
    dydt = 1
    cvx = 2
    map = 3
    traj = 4
    nil = 5


class Integrators(Enum):
    rk45 = 0
    lsoda = 1
    dop853 = 2
    implicit_midpoint = 3
    rk4 = 4
    euler = 5


# --------------------------------------------------------------------------------------------
# fixed-step solver objects (protocol: gds/ode/midpoint.py:10-32, itself a mirror of scipy OdeSolver)
# --------------------------------------------------------------------------------------------
STEP_MERGE_RTOL = 1e-9  # a remainder within (1+rtol)*max_step is taken as ONE step ending exactly at t_bound


class _FixedStep:
    """attrs t, y, fun, t_bound, status, max_step, n and step(), as fds expects (fds.py:157,171,196,285-288,362-363)."""

    def __init__(self, fun, t0, y0, t_bound, max_step=1e-3):
        self.t = t0
        self.y = y0  # aliases y0 exactly like midpoint.py:13 (fds.set_initial writes through it)
        self.fun = fun
        self.t_bound = t_bound
        self.status = 'running' if t0 < t_bound else 'finished'
        self.max_step = max_step
        self.n = y0.size

    def _advance(self, t, y, h):
        raise NotImplementedError

    def step(self):
        if self.status != 'running':
            raise RuntimeError('Attempt to step on a failed or finished solver.')
        if self.n == 0 or self.t >= self.t_bound:
            self.status = 'finished'
            return
        rem = self.t_bound - self.t
        if rem <= self.max_step * (1.0 + STEP_MERGE_RTOL):
            h, t_new = rem, self.t_bound
        else:
            h, t_new = self.max_step, self.t + self.max_step
        self.y = self._advance(self.t, self.y, h)
        self.t = t_new
        if self.t >= self.t_bound:
            self.status = 'finished'


class RK4(_FixedStep):
    def _advance(self, t, y, h):
        k1 = self.fun(t, y)
        k2 = self.fun(t + 0.5 * h, y + (0.5 * h) * k1)
        k3 = self.fun(t + 0.5 * h, y + (0.5 * h) * k2)
        k4 = self.fun(t + h, y + h * k3)
        acc = k1 + 2.0 * k2
        acc = acc + 2.0 * k3
        acc = acc + k4
        return y + (h / 6.0) * acc


class Euler(_FixedStep):
    def _advance(self, t, y, h):
        return y + h * self.fun(t, y)


# --------------------------------------------------------------------------------------------
# small helpers (gds/utils/common.py:27-36,55-56,77-79; gds/utils/boundary.py:9-41)
# --------------------------------------------------------------------------------------------
def fun_ary(f):
    return len(inspect.signature(f).parameters)  # common.py:77-79


def dict_fun(m, def_val=None):
    return lambda x: m[x] if x in m else def_val  # common.py:55-56


class bidict(dict):
    """orientation-agnostic edge dictionary (common.py:27-36)."""

    def __getitem__(self, key):
        if dict.__contains__(self, key):
            return dict.__getitem__(self, key)
        return dict.__getitem__(self, (key[1], key[0]))

    def __contains__(self, key):
        return dict.__contains__(self, key) or dict.__contains__(self, (key[1], key[0]))


def const_edge_bc(dG, v):  # boundary.py:9-15
    def vel(e):
        if e in dG.edges or (e[1], e[0]) in dG.edges:
            return v
        return None
    return vel


def zero_edge_bc(dG):  # boundary.py:17-19
    return const_edge_bc(dG, 0.)


def combine_bcs(*bcs):  # boundary.py:21-41 -- static conditions take precedence over dynamic ones
    bcs = [dict_fun(bc) if isinstance(bc, dict) else bc for bc in bcs]
    static = [bc for bc in bcs if fun_ary(bc) == 1]
    dynamic = [bc for bc in bcs if fun_ary(bc) == 2]
    if dynamic:
        def fun(t, x):
            for bc in static:
                v = bc(x)
                if v is not None:
                    return v
            for bc in dynamic:
                v = bc(t, x)
                if v is not None:
                    return v
            return None
        return fun

    def fun(x):
        for bc in static:
            v = bc(x)
            if v is not None:
                return v
        return None
    return fun


def relu(a):
    return np.maximum(a, 0)  # common.py:105-107


# --------------------------------------------------------------------------------------------
# planar faces (gds/utils/graph/generators.py:444-511)
# --------------------------------------------------------------------------------------------
def _turn_angle(pos, a, b, c):
    """angle from direction a->b to direction b->c in (-pi, pi]  (generators.py:454-466)."""
    x1, y1 = pos[b][0] - pos[a][0], pos[b][1] - pos[a][1]
    x2, y2 = pos[c][0] - pos[b][0], pos[c][1] - pos[b][1]
    theta = np.arctan2(y2, x2) - np.arctan2(y1, x1)
    sign, magn = np.sign(theta), np.abs(theta) % (2 * np.pi)
    theta = (2 * np.pi - magn) if sign < 0 else magn
    return (theta - 2 * np.pi) if theta > np.pi else theta


def shoelace_area(pos, face):
    a = 0.0
    k = len(face)
    for i in range(k):
        x1, y1 = pos[face[i]]
        x2, y2 = pos[face[(i + 1) % k]]
        a += x1 * y2 - x2 * y1
    return 0.5 * a


def embedded_faces(G):
    """Faces of a positioned planar graph in the reference's discovery order; returns (faces, outer).

    Walk (generators.py:468-496): for every edge of G.edges(), both half-edges (u,v) then (v,u); an unseen
    half-edge (tail, head) starts a walk that always turns to the most counter-clockwise neighbour
    (largest turn angle, first maximum wins, never straight back to the tail); the face is the visited node
    tuple starting at `head`.  The reference then removes the first face whose polygon touches every node
    (shapely test, generators.py:498-506); on the lattices that is the unbounded face (checked against the
    reference on square/triangular/hexagonal lattices in tests/golden).
    """
    pos = nx.get_node_attributes(G, 'pos')
    if pos == {}:
        is_planar, _ = nx.check_planarity(G)
        if not is_planar:
            return [], None  # generators.py:531-534
        # generators.py:513-518: spring layout with fixed seed, then the positioned walk
        pos = nx.spring_layout(G, scale=0.9, center=(0, 0), iterations=500, seed=1, dim=2)
        nx.set_node_attributes(G, pos, 'pos')
    faces = []
    seen = set()
    for edge in G.edges():
        for half in (edge, (edge[1], edge[0])):
            if half in seen:
                continue
            tail, head = half
            path = [head]
            while True:
                best, best_angle = None, -float('inf')
                for node in G.neighbors(head):
                    if node != tail:
                        ang = _turn_angle(pos, tail, head, node)
                        if ang > best_angle:
                            best_angle, best = ang, node
                if best is None:
                    path = None
                    break
                if best == path[0]:
                    break
                path.append(best)
                tail, head = head, best
            if path is not None:
                faces.append(tuple(path))
                seen.add((path[-1], path[0]))
                for i in range(1, len(path)):
                    seen.add((path[i - 1], path[i]))
    # generators.py:498-506: the FIRST face whose closed polygon touches every node is taken as the outer face
    # (shapely Polygon.intersects(Point): orientation-agnostic, boundary counts).  A face whose bounding box does not
    # cover all nodes cannot qualify, which keeps the test cheap on lattices.
    xs = [pos[n][0] for n in G.nodes()]
    ys = [pos[n][1] for n in G.nodes()]
    bx0, bx1, by0, by1 = min(xs), max(xs), min(ys), max(ys)
    outer = None
    for i, f in enumerate(faces):
        fx = [pos[n][0] for n in f]
        fy = [pos[n][1] for n in f]
        if min(fx) > bx0 or max(fx) < bx1 or min(fy) > by0 or max(fy) < by1:
            continue
        poly = [(float(pos[n][0]), float(pos[n][1])) for n in f]
        if all(_closed_polygon_touches(poly, x, y) for x, y in zip(xs, ys)):
            outer = f
            del faces[i]
            break
    return faces, outer


def _closed_polygon_touches(pts, x, y):
    """point in or on a closed polygon (what shapely's Polygon.intersects(Point) answers)"""
    n = len(pts)
    inside = False
    for i in range(n):
        x1, y1 = pts[i]
        x2, y2 = pts[(i + 1) % n]
        cross = (x2 - x1) * (y - y1) - (y2 - y1) * (x - x1)
        if abs(cross) <= 1e-12 * max(1.0, abs(x2 - x1) + abs(y2 - y1)):
            if min(x1, x2) - 1e-12 <= x <= max(x1, x2) + 1e-12 and min(y1, y2) - 1e-12 <= y <= max(y1, y2) + 1e-12:
                return True
        if (y1 > y) != (y2 > y):
            xin = x1 + (y - y1) * (x2 - x1) / (y2 - y1)
            if xin > x:
                inside = not inside
    return inside


# --------------------------------------------------------------------------------------------
# lowering: nx.Graph -> index maps and sparse operators
# --------------------------------------------------------------------------------------------
class Lowered:
    """Index maps (gds.py:21-45) and incidence structure (gds.py:109) of a graph."""

    def __init__(self, G, need_faces=True):
        self.G = G
        self.nodes = {v: i for i, v in enumerate(G.nodes())}          # gds.py:21
        self.edges = bidict({e: i for i, e in enumerate(G.edges())})  # gds.py:23
        if hasattr(G, 'faces'):                                        # gds.py:31-32
            faces = G.faces
        elif need_faces:
            faces, self.outer_face = embedded_faces(G)                 # gds.py:34
        else:
            faces = []
        self.faces = {f: i for i, f in enumerate(faces)}               # gds.py:35
        V, E = len(self.nodes), len(self.edges)
        self.V, self.E, self.F = V, E, len(self.faces)
        self.edge_u = np.fromiter((self.nodes[e[0]] for e in self.edges), dtype=np.int64, count=E)
        self.edge_v = np.fromiter((self.nodes[e[1]] for e in self.edges), dtype=np.int64, count=E)
        # generators.py:600-607 + gds.py:39-40: weight attribute 'weight', default 1.0
        self.w = np.array([d.get('weight', 1.0) for (_, _, d) in G.edges(data=True)], dtype=float).reshape(E)
        self.omega = np.sqrt(self.w)
        ar = np.arange(E)
        # gds.py:109: B[u,e] = -sqrt(w_e), B[v,e] = +sqrt(w_e)  (nx.incidence_matrix(oriented=True) @ diag(sqrt w))
        self.B = sp.csr_matrix((np.concatenate([-self.omega, self.omega]),
                                (np.concatenate([self.edge_u, self.edge_v]), np.concatenate([ar, ar]))),
                               shape=(V, E))

    def curl_matrix(self):
        """B2 (F x E), gds.py:227-239: +sqrt(w_e) if the stored orientation of e follows the face traversal,
        -sqrt(w_e) if both endpoints lie on the face otherwise (note: endpoint test, not edge test)."""
        if self.F == 0:
            return sp.csr_matrix((1, self.E))  # gds.py:239
        edge_keys = list(self.edges.keys())
        inc = [[] for _ in range(self.V)]  # edges incident to a node
        for e, (a, b) in enumerate(zip(self.edge_u, self.edge_v)):
            inc[a].append(e)
            inc[b].append(e)
        rows, cols, vals = [], [], []
        for f, fi in self.faces.items():
            members = set(f)
            forward = {(f[i - 1], f[i]) for i in range(1, len(f))} | {(f[-1], f[0])}
            cand = sorted({e for n in f for e in inc[self.nodes[n]]})
            for e in cand:
                a, b = edge_keys[e]
                if a in members and b in members:
                    rows.append(fi)
                    cols.append(e)
                    vals.append(self.omega[e] if (a, b) in forward else -self.omega[e])
        return sp.csr_matrix((vals, (rows, cols)), shape=(self.F, self.E))

    def weighted_degree(self):
        """row sums of the weighted adjacency (gds/utils/graph/common.py:5-10)."""
        deg = np.zeros(self.V)
        np.add.at(deg, self.edge_u, self.w)
        np.add.at(deg, self.edge_v, self.w)
        return deg


# --------------------------------------------------------------------------------------------
# fds: evolution / initial / constraints / stepping  (gds/fds.py)
# --------------------------------------------------------------------------------------------
class fds:
    def __init__(self, X):
        self.X = X
        self.iX = {i: x for x, i in X.items()}
        self.ndim = len(X)
        self.iter_mode = IterationMode.none
        self._set_bcs()
        self.t0 = 0.
        self.y0_fun = lambda _: 0.

    # fds.py:30-135 (dydt and map branches; cvx/traj are out of scope)
    def set_evolution(self, dydt=None, order=1, max_step=1e-3, integrator=Integrators.rk4, solver_args={},
                      map_fun=None, dt=1.0, nil=False, lhs=None, refresh_cvx=True):
        assert sum(x is not None and x is not False for x in (dydt, map_fun, nil or None, lhs)) == 1, \
            'Exactly one evolution law must be specified'
        if dydt is not None:
            self.iter_mode = IterationMode.dydt
            self.dydt_fun = dydt
            self.max_step = max_step
            self._dt = max_step
            self.order = order
            self.solver_args = solver_args
            self.y0 = np.zeros(self.ndim * order)
            self.integrator_type = integrator
            self.integrator = make_integrator(integrator, self.dydt, self.t0, self.y0, max_step, solver_args)
        elif lhs is not None:   # fds.py:97-113
            self.iter_mode = IterationMode.cvx
            self.lhs_fun = lhs
            self.refresh_cvx = refresh_cvx
            self.initialized = False
            self.y0 = np.zeros(self.ndim)
            self._t = self.t0
            self._y = self.y0.copy()
        elif map_fun is not None:
            self.iter_mode = IterationMode.map
            self.map_fun = map_fun
            self.y0 = np.zeros(self.ndim)
            self._dt = dt
            self._t = self.t0
            self._n = self._t
            self._y = self.y0.copy()  # fix of fds.py:122 (`self.t0.copy()` on a float)
        else:
            self.iter_mode = IterationMode.nil
            self._t = 0
            self._y = np.zeros(self.ndim)
            self.y0 = np.zeros(self.ndim)

    # fds.py:137-173
    def set_initial(self, t0=0., y0=lambda _: 0.):
        assert self.iter_mode != IterationMode.none, 'Use set_evolution() before setting initial conditions'
        if isinstance(y0, np.ndarray):
            arr = y0
            y0 = lambda x: arr[self.X[x]]
        self.t0 = t0
        self.y0_fun = y0
        if self.iter_mode is IterationMode.dydt:
            self.integrator.t = t0
        elif self.iter_mode is IterationMode.map:
            self._t = t0
            self._n = int(t0)
        else:
            self._t = t0
        constrained = set(self.X_dirichlet)
        for x in self.X.keys():
            if x in constrained:
                continue
            i, v = self.X[x], y0(x)
            self.y0[i] = v
            if self.iter_mode is IterationMode.dydt:
                self.integrator.y[i] = v
            else:
                self._y[i] = v

    # fds.py:175-206
    def set_constraints(self, dirichlet={}, neumann={}, project=lambda x: x):
        assert self.iter_mode != IterationMode.none, 'Use set_evolution() before setting boundary conditions'
        self._set_bcs(dirichlet, neumann, project)
        self.y0[self.dirichlet_indices] = self.dirichlet_values
        self.y0 = project(self.y0)
        if self.iter_mode is IterationMode.dydt:
            # fds.py:196 -- negative offsets: for order>1 this addresses the LAST block (Appendix B.4)
            self.integrator.y[self.dirichlet_indices - self.ndim] = self.dirichlet_values
        elif self.iter_mode in (IterationMode.map, IterationMode.cvx):   # fds.py:197-203
            self._y = self.y0.copy()
        else:
            raise Exception('Unsupported iteration mode for set_constraints()')

    # fds.py:208-243
    def _set_bcs(self, dirichlet={}, neumann={}, project=lambda x: x):
        if isinstance(dirichlet, dict):
            dirichlet = dict_fun(dirichlet)
        if isinstance(neumann, dict):
            neumann = dict_fun(neumann)
        self.dirichlet_fun, self.neumann_fun, self.project_fun = dirichlet, neumann, project
        self.dynamic_dirichlet = fun_ary(dirichlet) > 1
        self.dynamic_neumann = fun_ary(neumann) > 1

        def populate(fun, dynamic):
            call = (lambda x: fun(0., x)) if dynamic else fun
            domain = [x for x in self.X if call(x) is not None]
            indices = np.array([self.X[x] for x in domain], dtype=np.intp)
            values = np.array([call(x) for x in domain], dtype=float).reshape(len(domain))
            return domain, indices, values

        self.X_dirichlet, self.dirichlet_indices, self.dirichlet_values = populate(dirichlet, self.dynamic_dirichlet)
        self.X_neumann, self.neumann_indices, self.neumann_values = populate(neumann, self.dynamic_neumann)
        both = set(self.X_dirichlet) & set(self.X_neumann)
        assert len(both) == 0, f'Dirichlet and Neumann conditions overlap on {both}'

    # fds.py:264-290
    def step(self, dt):
        if self.iter_mode is IterationMode.none:
            raise Exception('Evolution law not specified')
        elif self.iter_mode is IterationMode.dydt:
            self.integrator.t_bound = self.t + dt
            self.integrator.status = 'running'
            while self.integrator.status != 'finished':
                self.integrator.step()
                self.update_constraints(self.t)
                self.apply_constraints()
        elif self.iter_mode is IterationMode.map:
            self.step_map(dt)
        elif self.iter_mode is IterationMode.cvx:
            self.step_cvx(dt)
        else:
            self._t += dt

    # fds.py:317-328 -- minimise |lhs(t, y)|^2 subject to the Dirichlet values (CVXPY in the reference)
    def step_cvx(self, dt):
        self._t += dt
        if (not self.initialized) or self.refresh_cvx:
            self.update_constraints(self.t)
            n = self.ndim
            c = np.asarray(self.lhs_fun(self.t, np.zeros(n)), dtype=float).reshape(-1)
            A = np.empty((c.size, n))
            for j in range(n):          # lhs is affine in y: probe its linear part column by column
                e = np.zeros(n); e[j] = 1.0
                A[:, j] = np.asarray(self.lhs_fun(self.t, e), dtype=float).reshape(-1) - c
            fixed = np.zeros(n, dtype=bool)
            fixed[self.dirichlet_indices] = True
            y = np.zeros(n)
            y[self.dirichlet_indices] = self.dirichlet_values
            rhs = -(c + A[:, fixed] @ y[fixed])
            y[~fixed] = np.linalg.lstsq(A[:, ~fixed], rhs, rcond=None)[0]
            self._y = y
            self.apply_constraints()
            self.initialized = True

    # fds.py:292-305
    def dydt(self, t, y):
        self._dt = self.integrator.t - t
        self.update_constraints(t)
        n, order = self.ndim, self.order
        ret = np.zeros_like(y)
        for i in range(order - 1):
            ret[n * i:n * (i + 1)] = y[n * (i + 1):n * (i + 2)]
        diff = self.dydt_fun(t, y[n * (order - 1):]) if order > 1 else self.dydt_fun(t, y)
        diff = np.array(diff, dtype=float, copy=True).reshape(n)  # (the reference zeroes the user's array in place)
        diff[self.dirichlet_indices] = 0.
        ret[n * (order - 1):] = diff
        return ret

    # fds.py:332-339 with the documented (t, y) signature (Appendix B.1)
    def step_map(self, dt):
        self._t += dt
        if (self._t - self._n) >= self._dt:
            self._n = self._t
            self.update_constraints(self.t)
            self._y = np.array(self.map_fun(self.t, self.y), dtype=float)
            self.apply_constraints()
            self._y[self.dirichlet_indices] = self.dirichlet_values

    # fds.py:350-357
    def update_constraints(self, t):
        if self.dynamic_dirichlet:
            self.dirichlet_values = np.array([self.dirichlet_fun(t, x) for x in self.X_dirichlet], dtype=float)
        if self.dynamic_neumann:
            self.neumann_values = np.array([self.neumann_fun(t, x) for x in self.X_neumann], dtype=float)

    # fds.py:359-369
    def apply_constraints(self):
        if self.iter_mode is IterationMode.dydt:
            self.integrator.y[self.dirichlet_indices - self.ndim] = self.dirichlet_values
            self.integrator.y = self.project_fun(self.integrator.y)
        elif self.iter_mode is IterationMode.cvx:
            self._y = self.project_fun(self._y)   # boundary values are part of the solution (fds.py:364-366)
        elif self.iter_mode is IterationMode.map:
            self._y[self.dirichlet_indices] = self.dirichlet_values
            self._y = self.project_fun(self._y)

    # fds.py:373-400
    @property
    def y(self):
        if self.iter_mode is IterationMode.none:
            return np.zeros(self.ndim)
        if self.iter_mode is IterationMode.dydt:
            return self.integrator.y[:self.ndim]
        return self._y

    @property
    def t(self):  # fds.py:385-392 (no branch for a field without a law: None)
        if self.iter_mode is IterationMode.none:
            return None
        if self.iter_mode is IterationMode.dydt:
            return self.integrator.t
        return self._t

    @property
    def dt(self):
        return self._dt if self.iter_mode in (IterationMode.dydt, IterationMode.map) else 1e-3   # fds.py:395-400

    def __call__(self, x):
        return self.y[self.X[x]]

    def system(self, name):
        return System(self, {name: self})

    def reset(self):
        """fds.py:248-262: install the law again, then the initial and boundary conditions.  (The reference's dydt branch
        omits `integrator=` and so falls back to its default adaptive solver; the fixed-step scheme is kept here.)"""
        assert self.iter_mode != IterationMode.none
        dirichlet, neumann, project = self.dirichlet_fun, self.neumann_fun, self.project_fun
        if self.iter_mode is IterationMode.dydt:
            self.set_evolution(dydt=self.dydt_fun, order=self.order, max_step=self.max_step,
                               integrator=self.integrator_type, solver_args=self.solver_args)
        elif self.iter_mode is IterationMode.cvx:
            self.set_evolution(lhs=self.lhs_fun, refresh_cvx=self.refresh_cvx)
        elif self.iter_mode is IterationMode.map:
            self.set_evolution(map_fun=self.map_fun, dt=self._dt)
        else:
            self.set_evolution(nil=True)
        self._set_bcs()
        self.set_initial(t0=self.t0, y0=self.y0_fun)
        if self.iter_mode is not IterationMode.nil:
            self.set_constraints(dirichlet=dirichlet, neumann=neumann, project=project)


def make_integrator(kind, fun, t0, y0, max_step, solver_args):
    if kind == Integrators.rk4:
        return RK4(fun, t0, y0, np.inf, max_step=max_step, **solver_args)
    if kind == Integrators.euler:
        return Euler(fun, t0, y0, np.inf, max_step=max_step, **solver_args)
    raise ValueError(f'Unsupported integrator type: {kind}')  # fds.py:95


# --------------------------------------------------------------------------------------------
# fields on graph domains (gds/gds.py)
# --------------------------------------------------------------------------------------------
class gds(fds):
    def __init__(self, G, Gd, lowered=None):
        self.G, self.Gd = G, Gd
        self.low = lowered if lowered is not None else Lowered(G, need_faces=(Gd is not GraphDomain.nodes))
        low = self.low
        self.nodes, self.edges, self.faces = low.nodes, low.edges, low.faces
        self.edge_weights = low.w
        # gds.py:43
        self.edge_orientation = {**{e: 1 for e in self.edges}, **{(e[1], e[0]): -1 for e in self.edges}}
        X = {GraphDomain.nodes: self.nodes, GraphDomain.edges: self.edges, GraphDomain.faces: self.faces}[Gd]
        fds.__init__(self, X)
        self.incidence = low.B  # gds.py:109 (kept sparse)


class node_gds(gds):
    def __init__(self, G, lowered=None):
        gds.__init__(self, G, GraphDomain.nodes, lowered)
        B = self.incidence
        self.vertex_laplacian = (-B @ B.T).tocsr()      # gds.py:140
        self.dirichlet_laplacian = self.vertex_laplacian  # gds.py:141
        self.neumann_correction = np.zeros(self.ndim)     # gds.py:142

    def set_constraints(self, *args, **kwargs):
        fds.set_constraints(self, *args, **kwargs)
        # gds.py:116-120: zero the Dirichlet ROWS of L0; Neumann values enter as a constant source, set once
        keep = np.ones(self.ndim)
        keep[self.dirichlet_indices] = 0.
        self.dirichlet_laplacian = (sp.diags(keep) @ self.vertex_laplacian).tocsr()
        self.neumann_correction[self.neumann_indices] = self.neumann_values

    def partial(self, e):  # gds.py:149-150
        return np.sqrt(self.edge_weights[self.edges[e]]) * (self(e[1]) - self(e[0]))

    def grad(self, y=None):  # gds.py:152-154
        if y is None: y = self.y
        return self.incidence.T @ y

    def laplacian(self, y=None):  # gds.py:156-159
        if y is None: y = self.y
        return self.dirichlet_laplacian @ y + self.neumann_correction

    def bilaplacian(self, y=None):  # gds.py:161-164
        if y is None: y = self.y
        return self.dirichlet_laplacian @ self.laplacian(y)

    def advect(self, v_field, y=None):  # gds.py:166-177
        if isinstance(v_field, edge_gds):
            assert v_field.G is self.G, 'Incompatible domains'   # gds.py:168
            v_field = v_field.y
        if y is None: y = self.y
        B = self.incidence
        Bp = (-(B @ sp.diags(np.sign(v_field)))).maximum(0)   # relu(-B .* sign v)
        C = B @ sp.diags(np.asarray(v_field, dtype=float))    # B .* v
        return -(C @ (Bp.T @ y))

    def lie_advect(self, v_field, y=None):  # gds.py:179-189
        if isinstance(v_field, edge_gds):
            assert v_field.G is self.G, 'Incompatible domains'   # gds.py:184
            v_field = v_field.y
        if y is None: y = self.y
        B = self.incidence
        Bp = (B @ sp.diags(np.sign(v_field))).maximum(0) @ sp.diags(np.asarray(v_field, dtype=float))
        return Bp @ (B.T @ y)


class edge_gds(gds):
    def __init__(self, G, lowered=None):
        gds.__init__(self, G, GraphDomain.edges, lowered)
        B = self.incidence
        self.edge_laplacian = (-B.T @ B).tocsr()          # gds.py:197
        self.curl_face = self.low.curl_matrix().tocsr()    # gds.py:227-239
        # gds.py:253-262: dual weights |B|^T diag(1/(deg-1)) |B| with zero diagonal, inf -> 0
        deg = self.low.weighted_degree()
        with np.errstate(divide='ignore'):
            dinv = 1.0 / (deg - 1.0)
        dinv[dinv == np.inf] = 0
        aB = abs(B)
        dw = (aB.T @ sp.diags(dinv) @ aB).tolil()
        dw.setdiag(0)
        self.dual_weights = dw.tocsr()
        self.dual_weights.eliminate_zeros()

    def __call__(self, x):  # gds.py:273-274
        return self.edge_orientation[x] * self.y[self.X[x]]

    def div(self, y=None):  # gds.py:278-280
        if y is None: y = self.y
        return -(self.incidence @ y)

    def influx(self, y=None):  # gds.py:282-285
        if y is None: y = self.y
        return np.asarray((self.incidence @ sp.diags(np.asarray(y, dtype=float))).maximum(0).sum(axis=1)).ravel()

    def outflux(self, y=None):  # gds.py:287-290
        if y is None: y = self.y
        return np.asarray((-(self.incidence @ sp.diags(np.asarray(y, dtype=float)))).maximum(0).sum(axis=1)).ravel()

    def curl(self, y=None):  # gds.py:292-294
        if y is None: y = self.y
        return self.curl_face @ y

    def dd_(self, y=None, normal=None):  # gds.py:296-301
        if y is None: y = self.y
        ret = -(self.incidence.T @ (self.incidence @ y))
        ret[self.dirichlet_indices] = 0
        if normal is not None: ret[normal] = 0
        return ret

    def d_d(self, y=None):  # gds.py:303-307
        if y is None: y = self.y
        ret = -(self.curl_face.T @ (self.curl_face @ y))
        ret[self.dirichlet_indices] = 0
        return ret

    def laplacian(self, y=None, free=None, normal=None):  # gds.py:309-317
        ret = self.dd_(y=y, normal=normal) + self.d_d(y=y)
        if free is not None: ret[free] = 0
        return ret

    def bilaplacian(self, y=None):  # gds.py:319-324
        return self.laplacian(self.laplacian(y))

    def advect(self, y=None, weighted=True):  # gds.py:326-345
        if y is None: y = self.y
        y = np.asarray(y, dtype=float)
        s = np.sign(y)
        Bs = self.incidence @ sp.diags(s)
        Bm, Bp = (-Bs).maximum(0), Bs.maximum(0)
        F = Bm.T @ Bp - Bp.T @ Bm
        if weighted:
            F = F.multiply(self.dual_weights)
        F = sp.csr_matrix(F)
        Ft = F.T.tocsr()
        # A = diag(y) F^T' ... literally: (F^T .* y)^T .* s + (F^T .* s)^T .* y  with column broadcasting
        A = (Ft.multiply(y)).T.multiply(s) + (Ft.multiply(s)).T.multiply(y)
        return -(sp.csr_matrix(A) @ y)


    def advect_jac(self, y=None, single_precision_operators=True):
        """gds.py:372-395: Jacobian of advect_torch by autograd.  sign() and relu() have zero derivative almost
        everywhere, so F (built from s = sign y) is a constant of the differentiation and
            advect(y)_i = -sum_j F_ij (y_i s_j + s_i y_j) y_j
            J_ik = -( delta_ik sum_j F_ij |y_j| + y_i F_ik s_k + 2 s_i F_ik y_k ).
        The reference's torch copies of the operators are float32 (`incidence_torch`, `dual_weights_torch`,
        gds.py:270-271): with `single_precision_operators` the entries of B and of the dual weights are rounded the same
        way (the arithmetic itself stays float64, as torch promotes when y is float64)."""
        if y is None: y = self.y
        y = np.asarray(y, dtype=float)
        s = np.sign(y)
        B, dw = self.incidence, self.dual_weights
        if single_precision_operators:
            B = sp.csr_matrix(B.astype(np.float32).astype(np.float64))
            dw = sp.csr_matrix(dw.astype(np.float32).astype(np.float64))
        Bs = B @ sp.diags(s)
        Bm, Bp = (-Bs).maximum(0), Bs.maximum(0)
        F = np.asarray((Bm.T @ Bp - Bp.T @ Bm).multiply(dw).todense())
        J = -(np.diag(F @ np.abs(y)) + (y[:, None] * F) * s[None, :] + 2.0 * (s[:, None] * F) * y[None, :])
        return J

    def advect_jvp(self, v, y=None):
        """J(y) v in full float64 (no reference counterpart: the reference only forms the dense Jacobian)"""
        return self.advect_jac(y, single_precision_operators=False) @ np.asarray(v, dtype=float)

    def leray_project(self, y=None):  # gds.py:249,414-419 (the Dirichlet branch, gds.py:420-427, is not runnable)
        if y is None: y = self.y
        assert self.dirichlet_indices.size == 0
        B = self.incidence
        # the reference applies the dense projector I - B^T pinv(B B^T) B; same product without forming it
        phi = np.linalg.pinv((B @ B.T).toarray(), hermitian=True) @ (B @ np.asarray(y, dtype=float))
        return np.asarray(y, dtype=float) - B.T @ phi


class face_gds(gds):
    def __init__(self, G, lowered=None):
        gds.__init__(self, G, GraphDomain.faces, lowered)
        self.curl_face = self.low.curl_matrix().tocsr()  # gds.py:457-469

    def laplacian(self, y=None):  # gds.py:473-479
        if y is None: y = self.y
        return -(self.curl_face @ (self.curl_face.T @ y))


# --------------------------------------------------------------------------------------------
# coupling (gds/fds.py:408-546) and System (gds/system.py:16-57)
# --------------------------------------------------------------------------------------------
class coupled_fds:
    def __init__(self, *systems):
        assert len(systems) >= 1, 'Pass one or more systems to couple'
        assert all(s.t == 0. for s in systems), 'All systems must be at zero-time initial conditions.'
        assert all(s.iter_mode != IterationMode.none for s in systems), 'All systems must have evolution laws.'
        self.t0 = 0.
        self.cont = [s for s in systems if s.iter_mode is IterationMode.dydt]
        self.cvxs = [s for s in systems if s.iter_mode is IterationMode.cvx]
        self.maps = [s for s in systems if s.iter_mode is IterationMode.map]
        self.discrete_t = self.t0
        self.discrete_y = {id(s): s.y0 for s in self.cvxs + self.maps}
        self.has_integrator = len(self.cont) > 0
        if self.has_integrator:
            kind = self.cont[0].integrator_type
            assert all(s.integrator_type is kind for s in self.cont), \
                'All continuous dynamics must use the same integrator.'
            self.dydt_max_step = min(s.max_step for s in self.cont)      # fds.py:437
            self.dydt_y0 = np.concatenate([s.y0 for s in self.cont])     # fds.py:439
            args = {}
            for s in self.cont:
                args.update(s.solver_args)
            self.integrator = make_integrator(kind, self.dydt, self.t0, self.dydt_y0, self.dydt_max_step, args)
        last = 0
        for s in self.cont:  # fds.py:456-460: state views; y and t are re-bound to the shared integrator
            s.view = slice(last, last + s.y0.size)
            last += s.y0.size
            _rebind(s, y=lambda s_: self.integrator.y[s_.view], t=lambda _: self.t)
        for s in self.cvxs + self.maps:
            _rebind(s, y=lambda s_: self.discrete_y[id(s_)], t=lambda _: self.t)

    def step(self, dt):  # fds.py:471-489
        if not self.has_integrator:
            return self.step_discrete(dt)
        self.integrator.t_bound = self.t + dt
        self.integrator.status = 'running'
        while self.integrator.status != 'finished':
            self.integrator.step()
            if self.integrator.t > self.discrete_t:
                self.step_discrete(self.integrator.t - self.discrete_t)
            for s in self.cont:
                s.update_constraints(self.integrator.t)
                self.integrator.y[s.view][s.dirichlet_indices - s.ndim] = s.dirichlet_values
                self.integrator.y[s.view] = s.project_fun(self.integrator.y[s.view])

    def step_discrete(self, dt):  # fds.py:491-503
        for s in self.cvxs + self.maps:
            fds.step(s, dt)
        for s in self.cvxs + self.maps:
            np.copyto(self.discrete_y[id(s)], s._y)
        self.discrete_t += dt

    def dydt(self, t, y):  # fds.py:505-507
        self.step_discrete(t - self.discrete_t)
        return np.concatenate([s.dydt(t, y[s.view]) for s in self.cont])

    def reset(self):  # fds.py:509-527 (discrete_t is not reset there either)
        if self.has_integrator:
            args = {}
            for s in self.cont:
                args.update(s.solver_args)
            self.integrator = make_integrator(self.cont[0].integrator_type, self.dydt, self.t0, self.dydt_y0,
                                              self.dydt_max_step, args)
        for s in self.cvxs + self.maps:
            s.reset()
            self.discrete_y[id(s)] = s._y.copy()

    def observables(self):
        return self.cont + self.cvxs + self.maps

    @property
    def t(self):
        return self.integrator.t if self.has_integrator else self.discrete_t


def _rebind(obj, **props):
    """common.py:46-52: give one instance its own property set through a throw-away subclass."""
    obj.__class__ = type(obj.__class__.__name__ + '_coupled', (obj.__class__,),
                         {k: property(v) for k, v in props.items()})


class System:
    def __init__(self, stepper, observables):
        self._stepper, self._observables = stepper, observables

    @property
    def stepper(self):
        return self._stepper

    @property
    def observables(self):
        return self._observables

    def step(self, dt):
        self._stepper.step(dt)

    @property
    def t(self):
        return self._stepper.t

    def solve(self, T, dt):
        """system.py:38-57, except that every observable is returned as an ndarray (Appendix B.8)."""
        data = {k: [v.y.copy()] for k, v in self._observables.items()}
        time, t = [0.], 0.
        while t < T:
            self._stepper.step(dt)
            for k, v in self._observables.items():
                data[k].append(v.y.copy())
            t += dt
            time.append(t)
        return np.array(time), {k: np.array(v) for k, v in data.items()}


def couple(observables):  # fds.py:542-546
    steppables = [o for o in observables.values() if isinstance(o, fds)]
    return System(coupled_fds(*steppables), observables)
