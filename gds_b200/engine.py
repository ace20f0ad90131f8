"""Device stepping engine: one concatenated state vector, one CUDA-graph-captured step.

Replaces the Python stepping loop of the reference -- fds.step_dydt / dydt / step_map / update_constraints /
apply_constraints (gds/fds.py:284-305,332-369) and coupled_fds.step_continuous / dydt (gds/fds.py:477-507) -- and
the solver object behind it (protocol of gds/ode/midpoint.py:10-32).  Host work per `advance()` is only: the step
table (t_i, h_i), the user's time-varying Dirichlet callables evaluated at the step-end times (exactly the values
the reference consumes, fds.py:289-290), and one C call that replays the captured step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import lib, check, ptr
from .device import Buffer, Context, Mask
from .trace import Lowering, Stage, Time, TraceError, as_expr, pass_desc, trace_scope
from .types import Integrators

STEP_MERGE_RTOL = 1e-9  # a remainder within (1+rtol)*max_step is taken as one step ending exactly at t_bound


def step_table(t, t_bound, max_step):
    """(t_i, h_i) of the fixed steps from t to t_bound: h = min(max_step, t_bound - t), the last step lands exactly
    on t_bound (solver protocol gds/ode/midpoint.py:20-32 with scipy's clamping of the final step)."""
    ts, hs = [], []
    while t < t_bound:
        rem = t_bound - t
        if rem <= max_step * (1.0 + STEP_MERGE_RTOL):
            h, t_new = rem, t_bound
        else:
            h, t_new = max_step, t + max_step
        ts.append(t)
        hs.append(h)
        t = t_new
    return np.asarray(ts, dtype=np.float64), np.asarray(hs, dtype=np.float64), t


def _reads_stage(e, seen=None):
    """does the expression DAG contain the stage state (`y` argument of the law) or the stage time?"""
    seen = set() if seen is None else seen
    if id(e) in seen:
        return False
    seen.add(id(e))
    if isinstance(e, (Stage, Time)):
        return True
    return any(_reads_stage(a, seen) for a in e.args)


def overlap_split(send_rows, n_own, max_fraction=0.25):
    """(lo_end, hi_begin), multiples of 128, such that [0, lo_end) and [hi_begin, n_own) hold every row in `send_rows`
    (the owned rows adjacent to a cut: they are sent to, and read ghost rows from, a neighbour) and are as short as
    possible -- the cut goes through the widest gap between boundary rows.  None when the two ranges would cover more than
    `max_fraction` of the owned rows (nothing to hide the exchange behind)."""
    r = np.unique(np.asarray(send_rows, dtype=np.int64))
    if r.size == 0 or n_own < 1024:
        return None
    lo = np.concatenate([[0], (r + 128) // 128 * 128])                     # lo_end if rows r[:i] go to the low range
    hi = np.concatenate([r // 128 * 128, [n_own // 128 * 128]])            # hi_begin if rows r[i:] go to the high range
    ok = hi >= lo
    if not ok.any():
        return None
    cost = np.where(ok, lo + (n_own - hi), np.iinfo(np.int64).max)
    i = int(np.argmin(cost))
    if cost[i] > max_fraction * n_own:
        return None
    return int(lo[i]), int(hi[i])


class StepEngine:
    """Compiled system of one or more fields advanced together."""

    def __init__(self, fields, scheme, max_step, t0=0.0, ctx: Context = None, record=False, allow_collapse=True):
        self.record = bool(record)   # snapshot kernel inside the captured step (System.solve)
        self.allow_collapse = bool(allow_collapse)
        assert len(fields) >= 1
        self.fields = list(fields)
        self.scheme = scheme
        self.max_step = float(max_step)
        self.t = float(t0)
        self.low = self.fields[0]._low
        for f in self.fields:
            if f._low is not self.low:
                raise ValueError('fields advanced together must live on the same graph object')
        self.ctx = ctx or self.low.ctx
        self.part = self.low.part          # this rank's part of a distributed graph (None: the whole graph is here)
        self.h = None
        self.offsets = {}
        off, self._gaps = 0, []
        for f in self.fields:
            start = (off + 3) // 4 * 4          # every field starts on a 32-byte boundary (128-bit loads / stores)
            if start > off:
                self._gaps.append((off, start - off))
            self.offsets[id(f)] = start
            off = start + f.ndim * f._order()
        self.n_state = off
        self.externals = {}      # id(field) -> [field, Buffer, version]
        self._keep = []
        self._read_cache = {}
        self.version = 0
        self.steps_taken = 0
        self._build()

    # ---- compilation -------------------------------------------------------------------------------------
    def _resolve_leaf(self, e):
        f = e.field
        if id(f) in self.offsets:
            off = self.offsets[id(f)]
            if isinstance(e, Stage):
                return (L.VEC_STAGE, None, off + f.ndim * (f._order() - 1))   # fds.py:300: last block is handed to dydt
            return (L.VEC_STATE, None, off)                                   # fds.py:378: y = integrator.y[:ndim]
        if isinstance(e, Stage):
            raise TraceError('stage state of a field outside this system')
        if getattr(f, '_dev_y', None) is not None:
            return (L.VEC_BUF, f._dev_y, 0)       # an `lhs=` field: its state already lives on the device
        if id(f) not in self.externals:
            self.externals[id(f)] = [f, Buffer(self.ctx, f.ndim), None]
        return (L.VEC_BUF, self.externals[id(f)][1], 0)

    def _build(self):
        scheme_code = {Integrators.rk4: L.SCHEME_RK4, Integrators.euler: L.SCHEME_EULER, 'map': L.SCHEME_MAP}[self.scheme]
        # run every law once with symbolic arguments (this may discover faces, which re-lowers the graph) ...
        roots = []
        for f in self.fields:
            with trace_scope():
                root = f.map_fun(Time(), Stage(f)) if self.scheme == 'map' else f.dydt_fun(Time(), Stage(f))
            root = as_expr(root, f._domain_code, f._low)
            if root.domain is not None and root.domain != f._domain_code:
                raise ValueError('evolution law returns a vector over the wrong domain')
            roots.append(root)
        # A law that reads accepted state only (`obs.y`, operators without an argument: the idiom of
        # gds/examples/fluid.py:33-34 and wave.py:14 -- SURVEY.md App. B.3) has the same derivative at all four RK4
        # stages: the step then runs every pass ONCE and combines the stages in registers (bit-identical to the staged
        # form, gds_system_set_collapsed).  Needs: no stage-state or stage-time leaf anywhere, fields of order <= 2.
        import os
        self.collapsed = (self.scheme == Integrators.rk4 and self.allow_collapse and not os.environ.get('GDS_B200_NO_COLLAPSE')
                          and all(f._order() <= 2 for f in self.fields)
                          and not any(_reads_stage(r) for r in roots))
        self.n_stages = 4 if (self.scheme == Integrators.rk4 and not self.collapsed) else 1
        # ... then bind to the lowered operators; the engine keeps them alive for as long as the captured step runs
        self.dev = self.low.dev
        h = L.c_vp()
        check(lib.gds_system_create(self.ctx.h, self.dev.h, self.n_state, scheme_code, C.byref(h)))
        self.h = h
        import weakref
        self._fin = weakref.finalize(self, lib.gds_system_destroy, h)
        if self.collapsed:
            check(lib.gds_system_set_collapsed(self.h, 1))
        if self.scheme == 'map' and self.part is None and not os.environ.get('GDS_B200_NO_SHADOW'):
            # y <- ... L f(y) ...: keep f(y) beside the state instead of evaluating f for every gathered neighbour
            check(lib.gds_system_set_shadow(self.h, 1))
        # (sub-expressions of accepted state are hoisted out of the stages into once-per-step passes; a collapsed step
        # has nothing to hoist them out of)
        lowering = Lowering(self.ctx, self._resolve_leaf, hoist=not self.collapsed,
                            single_hop=self.part is not None and not hasattr(self.part, 'gid_dom'))
        self.lowering = lowering
        self.passes = []
        stat_idx, stat_val, dyn_idx = [], [], []
        self.dyn_fields = []
        for f, root in zip(self.fields, roots):
            off, n, order = self.offsets[id(f)], f.ndim, f._order()
            dmask = None
            if self.scheme != 'map' and f.dirichlet_indices.size:
                dmask = Mask(self.ctx, n, f._low.rows_to_dev(f._domain_code, f.dirichlet_indices))   # fds.py:303
                self._keep.append(dmask)
            lowering.rhs_pass(root, f._domain_code, f._low, off + n * (order - 1), dmask)
            for p in lowering.passes:
                d, keep = pass_desc(p)
                check(lib.gds_system_add_pass(self.h, C.byref(d)))
            self.passes.extend(lowering.passes)      # kept for introspection (tests, kernel-kind queries)
            lowering.passes = []
            for i in range(order - 1):                                   # fds.py:297-298
                check(lib.gds_system_add_shift(self.h, off + n * i, off + n * (i + 1), n))
            # Dirichlet overwrite after every step: fds.py:362 addresses `dirichlet_indices - ndim`, i.e. the LAST
            # order block
            if f.dirichlet_indices.size:
                idx = off + n * (order - 1) + f._low.rows_to_dev(f._domain_code, f.dirichlet_indices)
                if f.dynamic_dirichlet:
                    dyn_idx.append(idx)
                    self.dyn_fields.append(f)
                else:
                    stat_idx.append(idx)
                    stat_val.append(np.asarray(f.dirichlet_values, dtype=np.float64))
        for (o, k) in self._gaps:
            check(lib.gds_system_add_hold(self.h, o, k))
        if stat_idx:
            i, v = L.as_i64(np.concatenate(stat_idx)), L.as_f64(np.concatenate(stat_val))
            check(lib.gds_system_set_dirichlet(self.h, ptr(i), ptr(v), i.size, 0))
        self.n_dyn = 0
        if dyn_idx:
            i = L.as_i64(np.concatenate(dyn_idx))
            self.n_dyn = i.size
            check(lib.gds_system_set_dirichlet(self.h, ptr(i), None, i.size, 1))
        self.p2p = False
        self.overlap = None
        self._stream = None
        if self.record:
            check(lib.gds_system_enable_recording(self.h))
        check(lib.gds_ctx_set_stream(self.ctx.h, None))         # the step is captured on the context's own stream
        if self.part is not None:
            self._attach_peers()
        check(lib.gds_system_finalize(self.h))
        kps, nbytes = L.c_i64(), L.c_i64()
        check(lib.gds_system_info(self.h, C.byref(kps), C.byref(nbytes)))
        self.kernels_per_step = kps.value
        m = [L.c_i32(), L.c_i32(), L.c_i32()]
        check(lib.gds_system_modes(self.h, *[C.byref(x) for x in m]))
        self.modes = {'collapsed': bool(m[0].value), 'shadow': bool(m[1].value), 'overlap': bool(m[2].value)}
        self.device_bytes = nbytes.value + lowering.temp_bytes
        self.push_state()
        if self.part is not None:
            self._setup_partitioned()

    # ---- state movement ------------------------------------------------------------------------------------
    def _blocks_to_dev(self, f, host):
        """state of field f (order blocks of ndim rows, public row order) -> device row order"""
        low, n = f._low, f.ndim
        if f._domain_code != L.NODES or low.i2p is None:
            return L.as_f64(host)
        a = np.asarray(host, dtype=np.float64)
        return np.ascontiguousarray(np.concatenate([a[i:i + n][low.i2p] for i in range(0, a.size, n)]))

    def _blocks_from_dev(self, f, dev):
        low, n = f._low, f.ndim
        if f._domain_code != L.NODES or low.p2i is None:
            return dev
        return np.concatenate([dev[i:i + n][low.p2i] for i in range(0, dev.size, n)])

    def push_state(self):
        """upload every field's host state (what `integrator.y` holds in the reference)"""
        self._bind_stream()
        for f in self.fields:
            a = self._blocks_to_dev(f, f._host_state)
            check(lib.gds_system_set_state(self.h, ptr(a), self.offsets[id(f)], a.size))
            f._state_on_device = True
        self.version += 1
        self._read_cache.clear()

    def pull_state(self):
        self._bind_stream()
        for f in self.fields:
            a = np.empty(f.ndim * f._order(), dtype=np.float64)
            check(lib.gds_system_get_state(self.h, ptr(a), self.offsets[id(f)], a.size))
            f._host_state = self._blocks_from_dev(f, a)
        self._check_p2p()

    def read(self, f, full=False):
        """current state slice of field f as a host array (cached until the next step)"""
        self._bind_stream()
        key = (id(f), full)
        if key not in self._read_cache:
            n = f.ndim * (f._order() if full else 1)
            a = np.empty(n, dtype=np.float64)
            check(lib.gds_system_get_state(self.h, ptr(a), self.offsets[id(f)], n))
            self._check_p2p()
            self._read_cache[key] = self._blocks_from_dev(f, a)
        return self._read_cache[key]

    def read_into(self, f, out):
        """copy the value block of field f into a caller-owned (ideally pinned) host array"""
        self._bind_stream()
        assert out.dtype == np.float64 and out.flags['C_CONTIGUOUS'] and out.size <= f.ndim
        if f._domain_code == L.NODES and f._low.p2i is not None:
            assert out.size == f.ndim
            tmp = np.empty(f.ndim, dtype=np.float64)
            check(lib.gds_system_get_state(self.h, ptr(tmp), self.offsets[id(f)], tmp.size))
            np.take(tmp, f._low.p2i, out=out)
            return out
        check(lib.gds_system_get_state(self.h, ptr(out), self.offsets[id(f)], out.size))
        self._check_p2p()
        return out

    def write(self, f, values, offset=0):
        self._bind_stream()
        a = L.as_f64(values)
        if f._domain_code == L.NODES and f._low.i2p is not None:
            assert offset % f.ndim == 0 and a.size % f.ndim == 0, 'a renumbered field is written block by block'
            a = self._blocks_to_dev(f, a)
        check(lib.gds_system_set_state(self.h, ptr(a), self.offsets[id(f)] + offset, a.size))
        if self.part is not None and self.h is not None and hasattr(self, 'hx'):
            self._exchange(self._state_ptr())      # ghost rows follow their owners
        self._read_cache.clear()
        self.version += 1

    def _refresh_externals(self):
        for rec in self.externals.values():
            f, buf, ver = rec
            v = f._state_version()
            if ver != v:
                buf.upload(f._low.to_dev(f._domain_code, f.y))
                rec[2] = v

    # ---- stepping ---------------------------------------------------------------------------------------------
    def _bc_table(self, t_end):
        """values of the time-varying Dirichlet conditions at the step-end times (fds.py:350-353,289-290)"""
        if self.n_dyn == 0:
            return None
        tab = np.empty((t_end.size, self.n_dyn), dtype=np.float64)
        col = 0
        for f in self.dyn_fields:
            k = f.dirichlet_indices.size
            for i, t in enumerate(t_end):
                tab[i, col:col + k] = f._dirichlet_at(t)
            f.dirichlet_values = tab[-1, col:col + k].copy()
            col += k
        return tab

    def _launch(self, ts, hs, t_after):
        self._bind_stream()
        self._check_p2p()
        bc = self._bc_table(t_after)
        if self.part is not None and not self.p2p:
            self._run_partitioned(ts, hs, bc)
        else:
            check(lib.gds_system_step(self.h, ts.size, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None))
        self.steps_taken += int(ts.size)
        self.version += 1
        self._read_cache.clear()

    def plan(self, dt):
        """host part of an advance: step table and time-varying Dirichlet values (user callables run here)"""
        ts, hs, t_new = step_table(self.t, self.t + dt, self.max_step)
        t_after = np.append(ts[1:], t_new)
        return ts, hs, t_after, t_new, self._bc_table(t_after) if ts.size else None

    def plan_steps(self, k, h):
        """exactly k steps of size h from the current time"""
        if self.scheme == 'map':      # k applications of the map at t + h, t + 2h, ... (fds.py:332-339)
            ts = self.t + h * (np.arange(k, dtype=np.float64) + 1.0)
            return ts, np.zeros(k, dtype=np.float64), ts, float(ts[-1]), self._bc_table(ts)
        ts = self.t + h * np.arange(k, dtype=np.float64)
        hs = np.full(k, h, dtype=np.float64)
        t_after = ts + h
        return ts, hs, t_after, float(t_after[-1]), self._bc_table(t_after)

    def _bind_stream(self):
        """engines of one context may use different streams (a partitioned engine shares a torch stream with its
        transport): select this engine's before enqueueing"""
        check(lib.gds_ctx_set_stream(self.ctx.h, self._stream.cuda_stream if self._stream is not None else None))

    def run(self, plan):
        """device part of an advance: replay the captured step once per table row (asynchronous)"""
        ts, hs, t_after, t_new, bc = plan
        if ts.size == 0:
            return
        self._bind_stream()
        self._check_p2p()
        if self.part is not None and not self.p2p:
            self._run_partitioned(ts, hs, bc)
        else:   # one GPU, or peer-to-peer halo fused into the captured step
            check(lib.gds_system_step(self.h, ts.size, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None))
        self.steps_taken += int(ts.size)
        self.version += 1
        self._read_cache.clear()
        self.t = t_new
        for f in self.fields:
            f._dt = float(hs[-1])

    def step_ensemble(self, members_in, members_out, dt):
        """advance every member of an ensemble of HOST-resident states from t to t + dt: the loop
        `set_initial(y0=member); step(dt); read y` of the reference (fds.py:137-162,264-290) as one call.  members_in /
        members_out: per member, one float64 array per field of the system (its state blocks, public row order; pinned
        memory lets the uploads, the steps and the read-backs of consecutive members overlap).  Bit-identical to
        write / run / read_into member by member.  Afterwards the engine holds the last member's state at t + dt."""
        import ctypes
        if self.part is not None:
            raise NotImplementedError('ensemble stepping on a partitioned engine (ghost rows need an exchange per upload)')
        if self.scheme == 'map' or any(not f._identity_projection() for f in self.fields):
            raise NotImplementedError('ensemble stepping covers continuous laws without a user projection')
        if len(members_in) != len(members_out):
            raise ValueError('one output per input member')
        nf = len(self.fields)
        for f in self.fields:
            if f._domain_code == L.NODES and f._low.i2p is not None:
                raise NotImplementedError('ensemble stepping of an internally renumbered field')
        for grp in list(members_in) + list(members_out):
            if len(grp) != nf:
                raise ValueError('one array per field of the system')
            for f, a in zip(self.fields, grp):
                if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
                        and a.size == f.ndim * f._order()):
                    raise ValueError('ensemble buffers are contiguous float64 arrays of ndim * order values per field')
        ts, hs, t_after, t_new, bc = self.plan(dt)
        M = len(members_in)
        if ts.size == 0 or M == 0:
            return
        self._refresh_externals()
        self._bind_stream()
        off = np.array([self.offsets[id(f)] for f in self.fields], dtype=np.int64)
        cnt = np.array([f.ndim * f._order() for f in self.fields], dtype=np.int64)
        PA = ctypes.c_void_p * (M * nf)
        pin = PA(*[a.ctypes.data for grp in members_in for a in grp])
        pout = PA(*[a.ctypes.data for grp in members_out for a in grp])
        check(lib.gds_system_step_ensemble(self.h, M, nf, ptr(off), ptr(cnt), pin, pout, ts.size, ptr(ts), ptr(hs),
                                           ptr(bc) if bc is not None else None))
        self.steps_taken += int(ts.size) * M
        self.version += 1
        self._read_cache.clear()
        self.t = t_new
        for f in self.fields:
            f._dt = float(hs[-1])

    # ---- one rank of a partitioned system ------------------------------------------------------
    def _halo_blocks(self):
        """(offset, domain) of every state block that has ghost rows"""
        deep = hasattr(self.part, 'gid_dom')
        for f in self.fields:
            if f._domain_code != L.NODES and not deep:
                raise NotImplementedError('edge and face fields need distribute(G, hops=k) with k >= 2')
        return [(self.offsets[id(f)] + f.ndim * i, f._domain_code) for f in self.fields for i in range(f._order())]

    def _check_ghost_layers(self):
        """Deep partitions recompute intermediates on the ghost layers instead of exchanging them: every gather eats
        one layer (an edge <- face gather floor(k/2) of them, k = largest face).  Propagate the level up to which each
        vector is valid through the lowered passes and make sure every law is still valid on the rows this rank owns."""
        part, h = self.part, self.part.hops
        kf = part.max_face // 2
        INF = 10 ** 9
        lv = {}

        def level_of(ref):
            kind, buf, _ = ref
            return lv.get(id(buf), h) if kind == L.VEC_BUF else h
        need = {L.NODES: 0, L.EDGES: 1, L.FACES: kf}
        OPN = {v: k for k, v in L.OP.items()}
        for p in self.passes:
            st = []
            for ins in p.code:
                op, a, b = OPN[ins & 0xff], (ins >> 8) & 0xff, (ins >> 16) & 0xff
                if op == 'PUSH_VEC':
                    st.append(level_of(p.vec[a]))
                elif op in ('PUSH_IMM', 'PUSH_T', 'PUSH_H'):
                    st.append(INF)
                elif op in ('N_LAP', 'N_BDOT', 'N_INFLUX', 'N_OUTFLUX'):
                    st.append(min(h - 1, level_of(p.vec[a]) - 1))
                elif op in ('N_ADVECT', 'N_LIE'):
                    st.append(min(h - 1, level_of(p.vec[a]) - 1, level_of(p.vec[b]) - 1))
                elif op == 'N_EAGG':
                    lv[id(p.out[b][0])] = min(h - 1, level_of(p.vec[a]) - 1)
                elif op == 'N_EAGGT':
                    lv[id(p.out[b][0])] = min(h - 1, level_of(p.vec[a]) - 1, level_of(p.vec[ins >> 24]) - 1)
                elif op in ('E_GRADT', 'F_CURL'):
                    st.append(level_of(p.vec[a]))
                elif op == 'E_CURLT':
                    st.append(min(h, level_of(p.vec[a])) - kf)
                elif op in ('E_ADVFIN', 'E_ADVS1'):
                    st.append(min(level_of(p.vec[a]), level_of(p.vec[b])))
                elif op == 'STORE':
                    lv[id(p.out[a][0])] = st[-1]
                elif op in ('ADD', 'SUB', 'MUL', 'DIV', 'POW', 'MAX', 'MIN', 'ATAN2'):
                    y = st.pop(); st[-1] = min(st[-1], y)
                elif op == 'DUP':
                    st.append(st[-1])
                elif op == 'SWAP':
                    st[-1], st[-2] = st[-2], st[-1]
                elif op == 'POP':
                    st.pop()
                elif op == 'GETL':
                    st.append(0)     # locals are not produced by the tracer; be conservative
            if p.epilogue == L.EPI_RHS and st and st[-1] < need[p.domain]:
                raise TraceError(f'this law needs more ghost layers than distribute(G, hops={h}) provides '
                                 f'(valid up to level {st[-1]}, needed {need[p.domain]})')

    def _attach_peers(self):
        """Peer-to-peer halo (GDS_B200_HALO=p2p, the default on an NCCL process group): the ranks swap the CUDA-IPC
        handles of their state buffers and the positions of their ghost rows, and the library fuses a push / wait pair
        into the captured step after every stage -- boundary values cross NVLink by direct stores into the neighbour's
        memory, with no host and no NCCL on the data path.  Otherwise (gloo test transport, GDS_B200_HALO=nccl) the
        host drives pack -> send/recv -> unpack between the stages.
        The transport is agreed on COLLECTIVELY: every rank takes part in the one all_gather below, and the fused halo
        is used only if every rank can use it (at most 8 neighbours each, same GDS_B200_HALO)."""
        import os
        import torch.distributed as dist
        part = self.part
        mode = os.environ.get('GDS_B200_HALO', '')
        if not (dist.is_available() and dist.is_initialized() and part.world > 1):
            return
        if not mode:
            mode = 'p2p' if dist.get_backend() == 'nccl' else 'host'
        want = mode == 'p2p' and len(part.peers) <= 8 and part.world <= 64
        deep = hasattr(part, 'gid_dom')
        blocks = self._halo_blocks()
        mine = {'want': want}
        if want:
            handles = (C.c_uint8 * 320)()
            check(lib.gds_system_ipc_export(self.h, handles))
            mine['handles'] = bytes(handles)
            mine['recv_from'] = {q: np.concatenate([off + part.recv_dom[d][q] for off, d in blocks]).astype(np.int64)
                                 for q in part.peers}
        everyone = [None] * part.world
        dist.all_gather_object(everyone, mine)
        if not all(e['want'] for e in everyone):
            return
        peers = [q for q in part.peers]
        src, dst, ptr_, hbytes = [], [], [0], b''
        for q in peers:
            s_idx = np.concatenate([off + part.send_dom[d][q] for off, d in blocks]).astype(np.int64)
            d_idx = np.asarray(everyone[q]['recv_from'][part.rank], dtype=np.int64)
            assert s_idx.size == d_idx.size, 'halo lists of two neighbouring ranks disagree'
            src.append(s_idx); dst.append(d_idx); ptr_.append(ptr_[-1] + s_idx.size)
            hbytes += everyone[q]['handles']
        src = L.as_i64(np.concatenate(src)) if src else np.zeros(0, dtype=np.int64)
        dst = L.as_i64(np.concatenate(dst)) if dst else np.zeros(0, dtype=np.int64)
        ptr_arr = L.as_i64(np.asarray(ptr_))
        pr = L.as_i32(np.asarray(peers, dtype=np.int32))
        hb = (C.c_uint8 * max(1, len(hbytes))).from_buffer_copy(hbytes or b'\0')
        # owned rows are a prefix of every block only in the node-only partitions; deep ones let the passes run over
        # the ghost rows (the push handshake keeps what the owners send)
        check(lib.gds_system_p2p_attach(self.h, part.rank, len(peers), ptr(pr), hb, ptr(ptr_arr), ptr(src), ptr(dst),
                                        -1 if deep else part.n_own))
        self.p2p = True
        self.overlap = None
        if not deep and peers and os.environ.get('GDS_B200_OVERLAP', '1') != '0':
            split = overlap_split(np.concatenate([part.send_rows[q] for q in peers]), part.n_own)
            if split is not None:
                check(lib.gds_system_p2p_set_overlap(self.h, split[0], split[1]))
                self.overlap = split

    def p2p_status(self):
        """(exchanges completed, error flag) of the fused peer-to-peer halo (synchronises the stream)"""
        a, b = L.c_i64(), L.c_i64()
        check(lib.gds_system_p2p_status(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def _check_p2p(self):
        """raise if an exchange of the fused halo has timed out (read from mapped host memory: no synchronisation, so
        it is looked at before every enqueue and after every read-back)"""
        if not self.p2p:
            return
        e = L.c_i64()
        check(lib.gds_system_p2p_error(self.h, C.byref(e)))
        if e.value:
            k = e.value - 1 if e.value < 100 else e.value - 101
            who = self.part.peers[k] if 0 <= k < len(self.part.peers) else '?'
            raise L.GdsError(f'peer-to-peer halo exchange timed out waiting for rank {who} (code {e.value}): a peer '
                             f'stopped or the ranks enqueued different numbers of steps; the state of rank '
                             f'{self.part.rank} is not valid any more')

    def _setup_partitioned(self):
        """halo lists over every state block, exchange buffers, and the transport stream"""
        import torch
        from .partition import HaloExchange
        part = self.part
        blocks = self._halo_blocks()
        self.hx = HaloExchange(part, blocks)
        if hasattr(part, 'gid_dom'):
            self._check_ghost_layers()
        self._halo = []
        for which in ('send', 'recv'):
            idx = L.as_i64(self.hx.state_indices(blocks, which))
            h = L.c_vp()
            check(lib.gds_halo_create(self.ctx.h, ptr(idx), idx.size, C.byref(h)))
            self._halo.append(h)
            import weakref
            weakref.finalize(self, lib.gds_halo_destroy, h)
        dev = torch.device('cuda', self.ctx.device)
        self._send = torch.empty(max(1, self.hx.n_send), dtype=torch.float64, device=dev)
        self._recv = torch.empty(max(1, self.hx.n_recv), dtype=torch.float64, device=dev)
        # kernels and the transfers are ordered through one torch-visible stream (made current around every exchange;
        # not the legacy default stream, whose handle 0 means "the context's own stream" to gds_ctx_set_stream)
        self._torch = torch
        self._stream = torch.cuda.Stream(dev)
        check(lib.gds_ctx_set_stream(self.ctx.h, self._stream.cuda_stream))
        self.exchanges = 0
        self._exchange(self._state_ptr())

    def _state_ptr(self):
        p = L.c_vp()
        check(lib.gds_system_state_devptr(self.h, C.byref(p)))
        return p

    def _exchange(self, vec_ptr):
        """refresh the ghost rows of one state-shaped device vector from their owners"""
        if not self.part.peers:
            return
        check(lib.gds_halo_pack(self._halo[0], vec_ptr, self._send.data_ptr()))
        with self._torch.cuda.stream(self._stream):
            for r in self.hx.exchange(self._send, self._recv):
                r.wait()
        check(lib.gds_halo_unpack(self._halo[1], self._recv.data_ptr(), vec_ptr))
        self.exchanges += 1

    def _run_partitioned(self, ts, hs, bc):
        done = 0
        while done < ts.size:
            m = min(ts.size - done, 4096)
            check(lib.gds_system_begin_steps(self.h, m, ptr(ts[done:done + m]), ptr(hs[done:done + m]),
                                             ptr(bc[done:done + m]) if bc is not None else None))
            out = L.c_vp()
            for _ in range(m):
                for st in range(self.n_stages):
                    check(lib.gds_system_run_stage(self.h, st, 0, -1))
                    if st < self.n_stages - 1:
                        check(lib.gds_system_stage_output(self.h, st, C.byref(out)))
                        self._exchange(out)
                check(lib.gds_system_end_step(self.h))
                self._exchange(self._state_ptr())
            done += m

    def solve_record(self, n_calls, dt, rec_calls, max_bytes=4 << 30):
        """`n_calls` consecutive step(dt) calls with the state after the calls listed in `rec_calls` (0-based, sorted)
        recorded ON THE DEVICE inside the captured step; returns [len(rec_calls), n_state] after one read-back per chunk
        (System.solve, gds/system.py:38-57, copies every observable to the host after every step)"""
        assert self.record and self.scheme != 'map' and self.part is None
        self._bind_stream()
        self._refresh_externals()
        rec_set = {int(c): i for i, c in enumerate(rec_calls)}
        out = np.empty((len(rec_calls), self.n_state), dtype=np.float64)
        per_chunk = max(1, int(max_bytes // (8 * self.n_state)))
        call = 0
        while call < n_calls:
            ts_l, hs_l, slots, t = [], [], [], self.t
            rows, first_row = 0, len([c for c in rec_calls if c < call])
            while call < n_calls and rows < per_chunk and len(ts_l) < 100000:
                ts, hs, t_new = step_table(t, t + dt, self.max_step)
                sl = np.full(ts.size, -1, dtype=np.int64)
                if call in rec_set and ts.size:
                    sl[-1] = rows
                    rows += 1
                ts_l.append(ts); hs_l.append(hs); slots.append(sl)
                t = t_new
                call += 1
            ts, hs, sl = np.concatenate(ts_l), np.concatenate(hs_l), L.as_i64(np.concatenate(slots))
            t_after = np.append(ts[1:], t)
            bc = self._bc_table(t_after)
            if rows:
                traj = Buffer(self.ctx, rows * self.n_state)
                check(lib.gds_system_set_trajectory(self.h, traj.h))
                check(lib.gds_system_step_record(self.h, ts.size, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None, ptr(sl)))
                out[first_row:first_row + rows] = traj.download().reshape(rows, self.n_state)
            else:
                check(lib.gds_system_step(self.h, ts.size, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None))
            self.steps_taken += int(ts.size)
            self.t = t
            for f in self.fields:
                f._dt = float(hs[-1])
        self.version += 1
        self._read_cache.clear()
        return out

    def run_timed(self, plan):
        """like run(), but kernel by kernel with CUDA events around every launch: returns (ms, launches) per class
        [node passes, edge passes, face passes, per-step tail] (measurement aid, synchronises).  The fused peer-to-peer
        exchange runs between the event pairs; the host-driven halo cannot be interleaved here."""
        ts, hs, t_after, t_new, bc = plan
        if self.part is not None and not self.p2p:
            raise NotImplementedError('timed stepping of a partitioned engine needs the peer-to-peer halo')
        self._bind_stream()
        self._check_p2p()
        ms, cnt = np.zeros(4, dtype=np.float64), np.zeros(4, dtype=np.int64)
        check(lib.gds_system_step_timed(self.h, ts.size, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None,
                                        ptr(ms), ptr(cnt)))
        self.steps_taken += int(ts.size)
        self.version += 1
        self._read_cache.clear()
        self.t = t_new
        return ms, cnt

    def advance(self, dt):
        """drive the system to t + dt in fixed steps of at most max_step (fds.py:284-290)"""
        self._refresh_externals()
        projecting = [f for f in self.fields if not f._identity_projection()]
        if not projecting:
            self.run(self.plan(dt))
            return
        # a user projection is arbitrary host code (fds.py:363): one device step, then a host round trip
        ts, hs, t_new = step_table(self.t, self.t + dt, self.max_step)
        t_after = np.append(ts[1:], t_new)
        for i in range(ts.size):
            self._launch(ts[i:i + 1], hs[i:i + 1], t_after[i:i + 1])
            for f in projecting:
                full = self.read(f, full=True)
                self.write(f, np.asarray(f.project_fun(full.copy()), dtype=np.float64))
            self.t = float(t_after[i])
        for f in self.fields:
            if ts.size:
                f._dt = float(hs[-1])

    def run_step_staged(self, t, h, t_after, before_stage):
        """ONE step enqueued stage by stage, `before_stage(s)` called before the passes of stage s are enqueued (a
        recurrence map coupled with this system may be applied there, fds.py:505-507)"""
        if self.part is not None:
            raise NotImplementedError('stage-by-stage stepping of a distributed system')
        ts, hs, ta = np.asarray([t], dtype=np.float64), np.asarray([h], dtype=np.float64), np.asarray([t_after], dtype=np.float64)
        bc = self._bc_table(ta)
        self._bind_stream()
        check(lib.gds_system_begin_steps(self.h, 1, ptr(ts), ptr(hs), ptr(bc) if bc is not None else None))
        for st in range(self.n_stages):
            before_stage(st)
            self._bind_stream()
            check(lib.gds_system_run_stage(self.h, st, 0, -1))
        check(lib.gds_system_end_step(self.h))
        self.steps_taken += 1
        self.version += 1
        self._read_cache.clear()
        self.t = float(t_after)
        for f in self.fields:
            f._dt = float(h)

    def apply_map(self, t_new, refresh=True):
        """one application of the recurrence map at time t_new (fds.py:332-339)"""
        if refresh:
            self._refresh_externals()
        ts, hs = np.asarray([t_new], dtype=np.float64), np.asarray([0.0], dtype=np.float64)
        self._launch(ts, hs, ts)
        self.t = t_new

    def sync(self):
        self.ctx.sync()
        self._check_p2p()

    def halo_check(self):
        """Partitioned engines: fetch, over the host-driven transport (NCCL send/recv), the rows this rank's neighbours
        own and compare them BIT FOR BIT with the ghost rows the stepping left here.  Returns the number of ghost values
        that differ on this rank (0 everywhere = every ghost row equals its owner's row)."""
        if self.part is None or not self.part.peers:
            return 0
        torch = self._torch
        self._bind_stream()
        state = self._state_ptr()
        have = torch.empty_like(self._recv)
        check(lib.gds_halo_pack(self._halo[1], state, have.data_ptr()))           # my ghost rows as they are
        check(lib.gds_halo_pack(self._halo[0], state, self._send.data_ptr()))     # my boundary rows for the neighbours
        with torch.cuda.stream(self._stream):
            for r in self.hx.exchange(self._send, self._recv):
                r.wait()
            self._stream.synchronize()
            bad = int((have.view(torch.int64) != self._recv.view(torch.int64)).sum().item()) if self.hx.n_recv else 0
        self._check_p2p()
        return bad
