"""The C-ABI shared library loads and exports every symbol include/gds_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import gds_b200
from gds_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'gds_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gds_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert len(names) >= 40
    lib = ctypes.CDLL(gds_b200.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/gds_b200.h but not exported'


def test_python_binding_covers_header():
    assert sorted(L.EXPORTED) == declared_symbols()


def test_abi_version_and_error_channel():
    assert L.lib.gds_abi_version() == 1
    h = L.c_vp()
    # a device that cannot exist: the call fails and the message is retrievable
    rc = L.lib.gds_ctx_create(10 ** 6, ctypes.byref(h))
    assert rc != 0
    assert len(L.lib.gds_last_error()) > 0


def test_host_only_context_never_computes():
    ctx = gds_b200.Context(-1)
    buf = gds_b200.Buffer(ctx, 8)
    try:
        buf.download()
    except gds_b200.GdsError as e:
        assert 'host-only' in str(e)
    else:
        raise AssertionError('download on a host-only context must fail')


def test_python_constants_equal_the_header():
    """opcodes, domains, schemes, operand kinds and limits mirrored in gds_b200/_lib.py are the header's values"""
    src = open(os.path.join(ROOT, 'include', 'gds_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    enums = {k: int(v) for k, v in re.findall(r'\b(GDS_[A-Z0-9_]+)\s*=\s*(-?\d+)', src)}
    defines = {k: int(v) for k, v in re.findall(r'#define\s+(GDS_[A-Z0-9_]+)\s+(\d+)\s*$', src, flags=re.M)}
    ops = {k[len('GDS_OP_'):]: v for k, v in enums.items() if k.startswith('GDS_OP_')}
    assert ops == dict(L.OP), (set(ops) ^ set(L.OP))
    for name in ('NODES', 'EDGES', 'FACES', 'SCHEME_EULER', 'SCHEME_RK4', 'SCHEME_MAP', 'WHEN_STAGE', 'WHEN_STEP',
                 'EPI_NONE', 'EPI_RHS', 'VEC_BUF', 'VEC_STATE', 'VEC_STAGE'):
        assert enums['GDS_' + name] == getattr(L, name), name
    for name in ('MAX_CODE', 'MAX_IMM', 'MAX_VEC', 'MAX_MASK', 'MAX_OUT'):
        assert defines['GDS_' + name] == getattr(L, name), name
    from gds_b200.trace import Lowering
    chain = {k[len('GDS_CHAIN_'):]: v for k, v in enums.items() if k.startswith('GDS_CHAIN_')}
    assert chain == Lowering.CHAIN


def test_new_device_entry_points_fail_loudly_without_a_device():
    """the multigrid hierarchy and the preconditioned solve have no CPU fallback: a host-only context refuses them"""
    ctx = gds_b200.Context(-1)
    h = L.c_vp()
    assert L.lib.gds_mg_create(ctx.h, ctypes.byref(h)) != 0
    assert b'host-only' in L.lib.gds_last_error()
    assert L.lib.gds_mg_create(None, ctypes.byref(h)) != 0
    assert L.lib.gds_mg_apply(None, None, None) != 0
    assert L.lib.gds_pcg_solve(ctx.h, None, None, None, None, None, None, None, None, None, 0, 1e-8, 10, 1, None, None) != 0
    assert L.lib.gds_system_step_ensemble(None, 1, 1, None, None, None, None, 1, None, None, None) != 0
    assert b'null system' in L.lib.gds_last_error()
