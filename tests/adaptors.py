"""API adaptors used by tests/cases.py (see its docstring)."""
import numpy as np


def _mk(mod):
    class API:
        node_gds = staticmethod(mod.node_gds)
        edge_gds = staticmethod(mod.edge_gds)
        face_gds = staticmethod(mod.face_gds)
        combine_bcs = staticmethod(mod.combine_bcs)
        zero_edge_bc = staticmethod(mod.zero_edge_bc)
        const_edge_bc = staticmethod(mod.const_edge_bc)
        couple = staticmethod(mod.couple)

        @staticmethod
        def evolve(obs, dydt, max_step, order=1, scheme='rk4'):
            kind = mod.Integrators.rk4 if scheme == 'rk4' else mod.Integrators.euler
            obs.set_evolution(dydt=dydt, order=order, max_step=max_step, integrator=kind)

        @staticmethod
        def evolve_lhs(obs, lhs):
            obs.set_evolution(lhs=lhs)

        @staticmethod
        def evolve_nil(obs):
            obs.set_evolution(nil=True)

        @staticmethod
        def evolve_map(obs, map_fun, dt=1.0):
            obs.set_evolution(map_fun=map_fun, dt=dt)
    return API


def oracle_api():
    from oracle import gds_oracle
    return _mk(gds_oracle)


def product_api():
    import gds_b200
    return _mk(gds_b200)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
