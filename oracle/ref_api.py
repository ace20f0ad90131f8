"""API adaptor that drives the UNMODIFIED reference (asrvsn/gds) with fixed-step solver objects -- TEST INFRASTRUCTURE.

Used by tests/golden/make_golden.py (golden vectors), tests/test_golden_fresh_cpu.py and the CPU legs of bench.py
(`--impl reference`, `cpu_baseline`).  The reference has no fixed-step explicit scheme (gds/types.py:47-51), so each field
is created with `Integrators.implicit_midpoint` and its `integrator` attribute is then re-pointed at an RK4 / Euler solver
object that follows the same protocol (gds/ode/midpoint.py:10-32) -- before set_initial / set_constraints, which write
into `integrator.y`.  Map mode is unreachable through the reference's `set_evolution` (fds.py:122 raises),
so the adaptor sets the attributes that branch would have set and lets the reference's own `step_map`
(fds.py:332-339) run.  The reference's numerical code runs as it is.
"""
import numpy as np

from .gds_oracle import RK4, Euler  # solver objects only
from .ref_harness import load_reference


def make_ref_api():
    ref = load_reference()

    class RefAPI:
        node_gds = staticmethod(ref.node_gds)
        edge_gds = staticmethod(ref.edge_gds)
        face_gds = staticmethod(ref.face_gds)
        combine_bcs = staticmethod(ref.combine_bcs)
        zero_edge_bc = staticmethod(ref.zero_edge_bc)
        const_edge_bc = staticmethod(ref.const_edge_bc)
    
        @staticmethod
        def evolve(obs, dydt, max_step, order=1, scheme='rk4'):
            obs.set_evolution(dydt=dydt, order=order, max_step=max_step, integrator=ref.Integrators.implicit_midpoint)
            cls = RK4 if scheme == 'rk4' else Euler
            obs.integrator = cls(obs.dydt, obs.t0, obs.y0, np.inf, max_step=max_step)
            obs._fixed_cls = cls
    
        @staticmethod
        def evolve_nil(obs):
            obs.set_evolution(nil=True)

        @staticmethod
        def evolve_lhs(obs, lhs):
            # the reference's own cvx path (fds.py:97-113); its `cvxpy` is oracle/cvx_shim.py where CVXPY is absent
            obs.set_evolution(lhs=lhs)
    
        @staticmethod
        def evolve_map(obs, map_fun, dt=1.0):
            # what fds.py:115-122 would set, with `_y = y0.copy()` in place of the failing `t0.copy()`
            obs.iter_mode = ref.IterationMode.map
            obs.map_fun = lambda y: map_fun(obs.t, y)   # step_map passes one argument (fds.py:337)
            obs.y0 = np.zeros(obs.ndim)
            obs._dt = dt
            obs._t = obs.t0
            obs._n = obs._t
            obs._y = obs.y0.copy()
    
        @staticmethod
        def couple(observables):
            sysobj = ref.couple(observables)
            st = sysobj.stepper
            if st.has_integrator:
                cls = st.systems[ref.IterationMode.dydt][0]._fixed_cls
                st.integrator = cls(st.dydt, st.t0, st.dydt_y0, np.inf, max_step=st.dydt_max_step)
            return sysobj
    return RefAPI
