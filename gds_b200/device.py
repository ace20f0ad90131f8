"""Thin Python handles over the C-ABI objects (context, device vectors, row masks, lowered graphs)."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib as L
from ._lib import lib, check, ptr


class Context:
    _default = {}

    def __init__(self, device: int = 0):
        h = L.c_vp()
        check(lib.gds_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, lib.gds_ctx_destroy, h)

    @classmethod
    def default(cls) -> 'Context':
        dev = int(os.environ.get('GDS_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
        if dev not in cls._default:
            cls._default[dev] = Context(dev)
        return cls._default[dev]

    def sync(self):
        check(lib.gds_ctx_sync(self.h))

    def timer_start(self):
        check(lib.gds_ctx_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = L.c_f64()
        check(lib.gds_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def counters(self):
        a, b = L.c_i64(), L.c_i64()
        check(lib.gds_ctx_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bandwidth_probe(self, n_read, n_write, n, reps=10) -> float:
        """GB/s of a gather-free kernel reading n_read and writing n_write float64 vectors of n elements"""
        g = L.c_f64()
        check(lib.gds_ctx_bandwidth_probe(self.h, int(n_read), int(n_write), int(n), int(reps), C.byref(g)))
        return g.value

    def flush_l2(self, nbytes=256 << 20):
        check(lib.gds_ctx_flush_l2(self.h, int(nbytes)))


class Buffer:
    """A float64 device vector."""

    def __init__(self, ctx: Context, n: int, data=None):
        h = L.c_vp()
        check(lib.gds_buf_create(ctx.h, int(n), C.byref(h)))
        self.h, self.n, self.ctx = h, int(n), ctx
        self._fin = weakref.finalize(self, lib.gds_buf_destroy, h)
        if data is not None:
            self.upload(data)

    def upload(self, data, offset=0):
        a = L.as_f64(data).ravel()
        check(lib.gds_buf_upload(self.h, int(offset), ptr(a), a.size))

    def download(self, offset=0, n=None) -> np.ndarray:
        n = self.n - offset if n is None else n
        out = np.empty(n, dtype=np.float64)
        check(lib.gds_buf_download(self.h, int(offset), ptr(out), n))
        return out

    def reduce(self, kind=0, offset=0, n=None) -> float:
        """sum (0) / max (1) / min (2) of the values, computed on the device"""
        r = L.c_f64()
        check(lib.gds_buf_reduce(self.h, int(offset), self.n - offset if n is None else int(n), int(kind), C.byref(r)))
        return r.value

    def fill(self, v, offset=0, n=None):
        check(lib.gds_buf_fill(self.h, int(offset), self.n - offset if n is None else int(n), float(v)))

    def devptr(self) -> int:
        p, n = L.c_vp(), L.c_i64()
        check(lib.gds_buf_devptr(self.h, C.byref(p), C.byref(n)))
        return p.value


class BufferView(Buffer):
    """non-owning view of device memory (a slice of a system's state) usable wherever a Buffer is"""

    def __init__(self, ctx: Context, dev_ptr: int, n: int):
        h = L.c_vp()
        check(lib.gds_buf_view(ctx.h, L.c_vp(dev_ptr), int(n), C.byref(h)))
        self.h, self.n, self.ctx = h, int(n), ctx
        self._fin = weakref.finalize(self, lib.gds_buf_destroy, h)


class Mask:
    """A set of rows as a device bitmap."""

    def __init__(self, ctx: Context, n_rows: int, rows):
        rows = L.as_i64(np.asarray(rows).ravel())
        h = L.c_vp()
        check(lib.gds_mask_create(ctx.h, int(n_rows), ptr(rows), rows.size, C.byref(h)))
        self.h, self.n_rows, self.n_set = h, int(n_rows), int(rows.size)
        self._fin = weakref.finalize(self, lib.gds_mask_destroy, h)


class DeviceGraph:
    """Lowered incidence structure on the device (gds_graph)."""

    def __init__(self, ctx: Context, n_nodes, edge_u, edge_v, weights=None, face_ptr=None, face_nodes=None,
                 hostgraph=None):
        h = L.c_vp()
        self.ctx = ctx
        if hostgraph is not None:
            w = None if weights is None else L.as_f64(weights)
            check(lib.gds_graph_create_from_host(ctx.h, hostgraph, ptr(w), C.byref(h)))
        else:
            eu, ev = L.as_i32(edge_u), L.as_i32(edge_v)
            w = None if weights is None else L.as_f64(weights)
            if face_ptr is None:
                fp, fn, nf = np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), 0
            else:
                fp, fn = L.as_i64(face_ptr), L.as_i32(face_nodes)
                nf = fp.size - 1
            check(lib.gds_graph_create(ctx.h, int(n_nodes), eu.size, ptr(eu), ptr(ev), ptr(w), nf, ptr(fp), ptr(fn),
                                       C.byref(h)))
        self.h = h
        self._fin = weakref.finalize(self, lib.gds_graph_destroy, h)
        a, b, c = L.c_i64(), L.c_i64(), L.c_i64()
        check(lib.gds_graph_sizes(h, C.byref(a), C.byref(b), C.byref(c)))
        self.V, self.E, self.F = a.value, b.value, c.value

    def rows(self, domain: int) -> int:
        return {L.NODES: self.V, L.EDGES: self.E, L.FACES: self.F}[domain]

    def device_bytes(self) -> int:
        b = L.c_i64()
        check(lib.gds_graph_device_bytes(self.h, C.byref(b)))
        return b.value

    def export_csr(self, kind: int):
        """(ptr, idx, sign) of B1 rows (0), B2 rows (1) or B2 columns (2): the device index maps, for parity checks."""
        nr, nnz = L.c_i64(), L.c_i64()
        check(lib.gds_graph_export_csr(self.h, kind, C.byref(nr), C.byref(nnz), None, None, None))
        p = np.empty(nr.value + 1, dtype=np.int64)
        i = np.empty(nnz.value, dtype=np.int32)
        s = np.empty(nnz.value, dtype=np.int32)
        check(lib.gds_graph_export_csr(self.h, kind, C.byref(nr), C.byref(nnz), ptr(p), ptr(i), ptr(s)))
        return p, i, s
