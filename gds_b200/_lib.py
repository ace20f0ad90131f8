"""ctypes binding of the C-ABI in include/gds_b200.h (libgds_b200.so, built in-tree by `make`).

There is no CPU fallback: importing this module without the built CUDA library raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GDS_B200_LIB: another build of the same library (kernel A/B runs); the default is the in-tree build
LIB_PATH = os.environ.get('GDS_B200_LIB') or os.path.join(_HERE, 'lib', 'libgds_b200.so')


class GdsError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(f'gds_b200: CUDA library not built ({LIB_PATH} missing). Run `make` at the repository root '
                      f'(or `python -c "import __graft_entry__ as g; g.build()"`). There is no CPU fallback.')

lib = C.CDLL(LIB_PATH)

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER

# opcodes / enums (mirror include/gds_b200.h)
NODES, EDGES, FACES = 0, 1, 3
SCHEME_EULER, SCHEME_RK4, SCHEME_MAP = 0, 1, 2
WHEN_STAGE, WHEN_STEP = 0, 1
EPI_NONE, EPI_RHS = 0, 1
VEC_BUF, VEC_STATE, VEC_STAGE = 0, 1, 2
OP = dict(
    END=0, PUSH_VEC=1, PUSH_IMM=2, PUSH_T=3, PUSH_H=4, DUP=5, SWAP=6, POP=7, SETL=8, GETL=9, STORE=10, MASK_ZERO=11,
    ADD=16, SUB=17, MUL=18, DIV=19, POW=20, MAX=21, MIN=22, ATAN2=23,
    NEG=32, ABS=33, SIGN=34, SQRT=35, EXP=36, LOG=37, SIN=38, COS=39, TAN=40, TANH=41, SINH=42, COSH=43, ATAN=44,
    SQUARE=45, RECIP=46, POWI=47,
    N_LAP=64, N_BDOT=65, N_INFLUX=66, N_OUTFLUX=67, N_ADVECT=68, N_LIE=69, N_EAGG=70, N_EAGGT=71,
    E_GRADT=80, E_CURLT=81, E_ADVFIN=82, E_ADVS1=83, F_CURL=96,
)
MAX_CODE, MAX_IMM, MAX_VEC, MAX_MASK, MAX_OUT, MAX_STACK = 96, 24, 16, 6, 4, 6


class VecRef(C.Structure):
    _fields_ = [('kind', c_i32), ('buf', c_vp), ('offset', c_i64)]


class PassDesc(C.Structure):
    _fields_ = [
        ('domain', c_i32), ('when', c_i32),
        ('n_code', c_i32), ('code', P(C.c_uint32)),
        ('n_imm', c_i32), ('imm', P(c_f64)),
        ('n_vec', c_i32), ('vec', P(VecRef)),
        ('n_mask', c_i32), ('mask', P(c_vp)),
        ('n_out', c_i32), ('out', P(c_vp)), ('out_offset', P(c_i64)),
        ('epilogue', c_i32), ('state_offset', c_i64), ('dirichlet', c_vp),
    ]


def _sig(name, *argtypes, restype=C.c_int):
    f = getattr(lib, name)
    f.argtypes = list(argtypes)
    f.restype = restype
    return f


_sig('gds_last_error', restype=C.c_char_p)
_sig('gds_abi_version')
_sig('gds_ctx_create', C.c_int, P(c_vp))
_sig('gds_ctx_destroy', c_vp)
_sig('gds_ctx_set_stream', c_vp, c_vp)
_sig('gds_ctx_sync', c_vp)
_sig('gds_ctx_timer_start', c_vp)
_sig('gds_ctx_timer_stop', c_vp, P(c_f64))
_sig('gds_ctx_counters', c_vp, P(c_i64), P(c_i64))
_sig('gds_ctx_flush_l2', c_vp, c_i64)
_sig('gds_ctx_bandwidth_probe', c_vp, c_i32, c_i32, c_i64, c_i32, P(c_f64))
_sig('gds_buf_create', c_vp, c_i64, P(c_vp))
_sig('gds_buf_view', c_vp, c_vp, c_i64, P(c_vp))
_sig('gds_buf_destroy', c_vp)
_sig('gds_buf_upload', c_vp, c_i64, c_vp, c_i64)
_sig('gds_buf_download', c_vp, c_i64, c_vp, c_i64)
_sig('gds_buf_fill', c_vp, c_i64, c_i64, c_f64)
_sig('gds_buf_devptr', c_vp, P(c_vp), P(c_i64))
_sig('gds_buf_reduce', c_vp, c_i64, c_i64, c_i32, P(c_f64))
_sig('gds_mask_create', c_vp, c_i64, c_vp, c_i64, P(c_vp))
_sig('gds_mask_destroy', c_vp)
_sig('gds_graph_create', c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, P(c_vp))
_sig('gds_graph_destroy', c_vp)
_sig('gds_graph_sizes', c_vp, P(c_i64), P(c_i64), P(c_i64))
_sig('gds_graph_export_csr', c_vp, c_i32, P(c_i64), P(c_i64), c_vp, c_vp, c_vp)
_sig('gds_graph_device_bytes', c_vp, P(c_i64))
_sig('gds_pass_run', c_vp, c_vp, P(PassDesc), c_f64)
_sig('gds_passes_run_cols', c_vp, c_vp, c_i32, P(PassDesc), c_f64, c_i64, c_i32, P(c_vp), P(c_i64))
_sig('gds_pass_kernel_kind', P(PassDesc), P(c_i32))
_sig('gds_cg_solve', c_vp, c_vp, P(PassDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_i32, c_i32, P(c_i32), P(c_f64))
_sig('gds_mg_aggregate', c_i64, c_vp, c_vp, c_vp, c_f64, c_vp, P(c_i64))
_sig('gds_mg_create', c_vp, P(c_vp))
_sig('gds_mg_destroy', c_vp)
_sig('gds_mg_add_level', c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp)
_sig('gds_mg_set_sweeps', c_vp, c_i32, c_i32)
_sig('gds_mg_info', c_vp, P(c_i32), P(c_i64), P(c_i64))
_sig('gds_mg_apply', c_vp, c_vp, c_vp)
_sig('gds_pcg_solve', c_vp, c_vp, P(PassDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_i32, c_i32, P(c_i32), P(c_f64))
_sig('gds_system_create', c_vp, c_vp, c_i64, c_i32, P(c_vp))
_sig('gds_system_destroy', c_vp)
_sig('gds_system_add_pass', c_vp, P(PassDesc))
_sig('gds_system_add_shift', c_vp, c_i64, c_i64, c_i64)
_sig('gds_system_add_hold', c_vp, c_i64, c_i64)
_sig('gds_system_set_dirichlet', c_vp, c_vp, c_vp, c_i64, c_i32)
_sig('gds_system_set_collapsed', c_vp, c_i32)
_sig('gds_system_set_shadow', c_vp, c_i32)
_sig('gds_system_finalize', c_vp)
_sig('gds_system_set_state', c_vp, c_vp, c_i64, c_i64)
_sig('gds_system_get_state', c_vp, c_vp, c_i64, c_i64)
_sig('gds_system_state_devptr', c_vp, P(c_vp))
_sig('gds_system_step', c_vp, c_i64, c_vp, c_vp, c_vp)
_sig('gds_system_step_ensemble', c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp)
_sig('gds_system_enable_recording', c_vp)
_sig('gds_system_set_trajectory', c_vp, c_vp)
_sig('gds_system_step_record', c_vp, c_i64, c_vp, c_vp, c_vp, c_vp)
_sig('gds_system_step_timed', c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp)
_sig('gds_system_begin_steps', c_vp, c_i64, c_vp, c_vp, c_vp)
_sig('gds_system_run_stage', c_vp, c_i32, c_i64, c_i64)
_sig('gds_system_stage_output', c_vp, c_i32, P(c_vp))
_sig('gds_system_end_step', c_vp)
_sig('gds_halo_create', c_vp, c_vp, c_i64, P(c_vp))
_sig('gds_halo_destroy', c_vp)
_sig('gds_halo_pack', c_vp, c_vp, c_vp)
_sig('gds_halo_unpack', c_vp, c_vp, c_vp)
_sig('gds_system_ipc_export', c_vp, c_vp)
_sig('gds_system_p2p_attach', c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64)
_sig('gds_system_p2p_status', c_vp, P(c_i64), P(c_i64))
_sig('gds_system_p2p_error', c_vp, P(c_i64))
_sig('gds_system_p2p_set_overlap', c_vp, c_i64, c_i64)
_sig('gds_system_info', c_vp, P(c_i64), P(c_i64))
_sig('gds_system_modes', c_vp, P(c_i32), P(c_i32), P(c_i32))
_sig('gds_hostgraph_lattice', c_i32, c_i64, c_i64, c_i64, c_f64, c_i32, P(c_vp))
_sig('gds_hostgraph_from_graph', c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, P(c_vp))
_sig('gds_hostgraph_sizes', c_vp, P(c_i64), P(c_i64), P(c_i64), P(c_i64), P(c_i32), P(c_i32))
_sig('gds_hostgraph_copy', c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp)
_sig('gds_hostgraph_destroy', c_vp)
_sig('gds_graph_create_from_host', c_vp, c_vp, c_vp, P(c_vp))

EXPORTED = [
    'gds_last_error', 'gds_abi_version', 'gds_ctx_create', 'gds_ctx_destroy', 'gds_ctx_set_stream', 'gds_ctx_sync',
    'gds_ctx_timer_start', 'gds_ctx_timer_stop', 'gds_ctx_counters', 'gds_ctx_flush_l2', 'gds_ctx_bandwidth_probe', 'gds_buf_create',
    'gds_buf_view', 'gds_buf_destroy', 'gds_buf_upload', 'gds_buf_download', 'gds_buf_fill', 'gds_buf_devptr', 'gds_buf_reduce', 'gds_mask_create',
    'gds_mask_destroy', 'gds_graph_create', 'gds_graph_destroy', 'gds_graph_sizes', 'gds_graph_export_csr',
    'gds_graph_device_bytes', 'gds_pass_run', 'gds_passes_run_cols', 'gds_pass_kernel_kind', 'gds_cg_solve', 'gds_mg_aggregate', 'gds_mg_create', 'gds_mg_destroy', 'gds_mg_add_level', 'gds_mg_set_sweeps', 'gds_mg_info', 'gds_mg_apply', 'gds_pcg_solve', 'gds_system_create', 'gds_system_destroy', 'gds_system_add_pass',
    'gds_system_add_shift', 'gds_system_add_hold', 'gds_system_set_dirichlet', 'gds_system_set_collapsed', 'gds_system_set_shadow', 'gds_system_finalize',
    'gds_system_set_state', 'gds_system_get_state', 'gds_system_state_devptr', 'gds_system_step', 'gds_system_step_ensemble', 'gds_system_step_timed', 'gds_system_info', 'gds_system_modes',
    'gds_system_begin_steps', 'gds_system_run_stage', 'gds_system_stage_output', 'gds_system_end_step',
    'gds_halo_create', 'gds_halo_destroy', 'gds_halo_pack', 'gds_halo_unpack',
    'gds_system_ipc_export', 'gds_system_p2p_attach', 'gds_system_p2p_status',
    'gds_system_p2p_error', 'gds_system_p2p_set_overlap',
    'gds_system_enable_recording', 'gds_system_set_trajectory', 'gds_system_step_record',
    'gds_hostgraph_lattice', 'gds_hostgraph_from_graph', 'gds_hostgraph_sizes', 'gds_hostgraph_copy',
    'gds_hostgraph_destroy', 'gds_graph_create_from_host',
]


def check(rc):
    if rc != 0:
        raise GdsError(lib.gds_last_error().decode('utf-8', 'replace'))


def ptr(arr):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if arr is None:
        return None
    assert arr.flags['C_CONTIGUOUS']
    return arr.ctypes.data_as(c_vp)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
