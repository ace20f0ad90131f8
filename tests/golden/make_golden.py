"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [case ...]

The reference is imported through oracle/ref_harness.py (import-time stand-ins for absent non-numerical
dependencies; numerical code untouched).  The reference has no fixed-step explicit scheme, so each field is
created with `Integrators.implicit_midpoint` and its `integrator` attribute is then re-pointed at an RK4/Euler
solver object that follows the same protocol (gds/ode/midpoint.py:10-32) -- before set_initial/set_constraints,
which write into `integrator.y`.  Map mode is unreachable through the reference's
`set_evolution` (fds.py:122 raises), so the adaptor sets the attributes that branch would have set and lets the
reference's own `step_map` (fds.py:332-339) run.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle.ref_harness import load_reference  # noqa: E402
from oracle.gds_oracle import RK4, Euler  # noqa: E402  (solver objects only)
import cases  # noqa: E402

warnings.simplefilter('ignore')
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


def main(argv):
    want = set(argv)
    for gname in cases.OP_GRAPHS:
        key = f'ops_{gname}'
        if want and key not in want:
            continue
        out = cases.operator_outputs(RefAPI, gname)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **out)
        print('wrote', key, {k: v.shape for k, v in out.items() if k.startswith('map') or k == 'n_faces'})
    for gname in cases.JAC_GRAPHS:
        key = f'jac_{gname}'
        if want and key not in want:
            continue
        out = cases.jacobian_outputs(RefAPI, gname)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **out)
        print('wrote', key, out['jac'].shape)
    for cname, fn in cases.QUIRK_CASES.items():
        key = f'quirk_{cname}'
        if want and key not in want:
            continue
        out = fn(RefAPI)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **{k: np.asarray(v) for k, v in out.items()})
        print('wrote', key, {k: np.shape(v) for k, v in out.items()})
    for gname, make in cases.RAGGED.items():
        key = f'ragged_{gname}'
        if want and key not in want:
            continue
        try:
            out = cases.run_ops(RefAPI, make())
        except Exception as exc:      # degenerate graphs the reference cannot construct (no edges, ...)
            print('reference cannot run', key, '-', type(exc).__name__, str(exc)[:80])
            continue
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **{k: np.asarray(v) for k, v in out.items()})
        print('wrote', key, sorted(out))
    for cname, fn in cases.API_CASES.items():
        key = f'api_{cname}'
        if want and key not in want:
            continue
        out = fn(RefAPI)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **{k: np.asarray(v) for k, v in out.items()})
        print('wrote', key, {k: np.shape(v) for k, v in out.items()})
    for cname, fn in cases.TRAJ_CASES.items():
        key = f'traj_{cname}'
        if want and key not in want:
            continue
        out = fn(RefAPI)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **out)
        print('wrote', key, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main(sys.argv[1:])
