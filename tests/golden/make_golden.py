"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [case ...]

The reference is imported through oracle/ref_harness.py (import-time stand-ins for absent non-numerical
dependencies; numerical code untouched).  The reference has no fixed-step explicit scheme, so each field is
created with `Integrators.implicit_midpoint` and its `integrator` attribute is then re-pointed at an RK4/Euler
solver object that follows the same protocol (gds/ode/midpoint.py:10-32) -- before set_initial/set_constraints,
which write into `integrator.y`.  Map mode is unreachable through the reference's
`set_evolution` (fds.py:122 raises), so the adaptor sets the attributes that branch would have set and lets the
reference's own `step_map` (fds.py:332-339) run.  `lhs=` fields (lhs_*.npz) run the reference's cvx path with
oracle/cvx_shim.py installed as `cvxpy` (CVXPY is absent): the reference assembles the programme, the stand-in solves it
exactly by dense least squares.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle.ref_api import make_ref_api  # noqa: E402
import cases  # noqa: E402

warnings.simplefilter('ignore')


RefAPI = make_ref_api()


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
    for cname, fn in cases.LHS_CASES.items():
        key = f'lhs_{cname}'
        if want and key not in want:
            continue
        out = fn(RefAPI)      # the reference's cvx path with oracle/cvx_shim.py as its `cvxpy` (exact least squares)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **out)
        print('wrote', key, {k: v.shape for k, v in out.items()})
    for cname, fn in cases.TRAJ_CASES.items():
        key = f'traj_{cname}'
        if want and key not in want:
            continue
        out = fn(RefAPI)
        np.savez_compressed(os.path.join(HERE, key + '.npz'), **out)
        print('wrote', key, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main(sys.argv[1:])
