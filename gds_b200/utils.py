"""Host-side helpers with the reference's semantics (gds/utils/common.py, gds/utils/boundary.py)."""
from __future__ import annotations

from inspect import signature
from typing import Any, Callable, Dict


class bidict(dict):
    """Edge dictionary whose lookups ignore orientation (gds/utils/common.py:27-36)."""

    def __getitem__(self, key):
        if dict.__contains__(self, key):
            return dict.__getitem__(self, key)
        return dict.__getitem__(self, (key[1], key[0]))

    def __contains__(self, key):
        return dict.__contains__(self, key) or dict.__contains__(self, (key[1], key[0]))


def fun_ary(f: Callable) -> int:
    """number of parameters of f -- decides static f(x) vs dynamic f(t, x) conditions (common.py:77-79)"""
    return len(signature(f).parameters)


def dict_fun(m: Dict, def_val: Any = None) -> Callable:
    return lambda x: m[x] if x in m else def_val  # common.py:55-56


def const_edge_bc(dG, v: float):
    """constant value along the edges of a boundary sub-graph, either orientation (boundary.py:9-15)"""
    def vel(e):
        if e in dG.edges or (e[1], e[0]) in dG.edges:
            return v
        return None
    return vel


def zero_edge_bc(dG):
    """no-slip condition on a boundary sub-graph (boundary.py:17-19)"""
    return const_edge_bc(dG, 0.)


def combine_bcs(*bcs):
    """merge dict / f(x) / f(t, x) conditions; static ones are consulted first (boundary.py:21-41)"""
    bcs = [dict_fun(bc) if isinstance(bc, dict) else bc for bc in bcs]
    static = [bc for bc in bcs if fun_ary(bc) == 1]
    dynamic = [bc for bc in bcs if fun_ary(bc) == 2]

    def first(fs, *a):
        for f in fs:
            v = f(*a)
            if v is not None:
                return v
        return None
    if dynamic:
        def fun(t, x):
            v = first(static, x)
            return v if v is not None else first(dynamic, t, x)
        return fun

    def fun(x):
        return first(static, x)
    return fun
