"""oracle/lsap_oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/lsap.c (the C restatement of the SciPy solver the reference
calls at detr/matcher.py:94) plus a pure-Python twin for tiny cases.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblsap_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/lsap.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "lsap.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/liblsap_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.lsap_oracle_f64.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int64, ctypes.c_int64, i64p, i64p]
        lib.lsap_oracle_f32.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int64, ctypes.c_int64, i64p, i64p]
        lib.lsap_oracle_batch_f32.argtypes = [ctypes.POINTER(ctypes.c_float), i64p, i64p, i64p, ctypes.c_int64, i64p, i64p, i64p]
        for f in (lib.lsap_oracle_f64, lib.lsap_oracle_f32, lib.lsap_oracle_batch_f32):
            f.restype = ctypes.c_int
        _lib = lib
    return _lib


def _raise(rc: int):
    # error texts follow SciPy's
    if rc == 1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == 2:
        raise ValueError("cost matrix is infeasible")
    if rc != 0:
        raise RuntimeError(f"lsap oracle failed rc={rc}")


def linear_sum_assignment(cost) -> tuple[np.ndarray, np.ndarray]:
    """Same contract as scipy.optimize.linear_sum_assignment for 2-D float input."""
    c = np.asarray(cost)
    if c.ndim != 2:
        raise ValueError("expected a matrix (2-D array)")
    nr, nc = c.shape
    n = min(nr, nc)
    rows = np.empty(n, dtype=np.int64)
    cols = np.empty(n, dtype=np.int64)
    lib = _load()
    i64p = ctypes.POINTER(ctypes.c_int64)
    if c.dtype == np.float32:
        c = np.ascontiguousarray(c)
        rc = lib.lsap_oracle_f32(c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nr, nc,
                                 rows.ctypes.data_as(i64p), cols.ctypes.data_as(i64p))
    else:
        c = np.ascontiguousarray(c, dtype=np.float64)
        rc = lib.lsap_oracle_f64(c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nr, nc,
                                 rows.ctypes.data_as(i64p), cols.ctypes.data_as(i64p))
    _raise(rc)
    return rows, cols


def linear_sum_assignment_batch(costs: list[np.ndarray]) -> list[tuple[np.ndarray, np.ndarray]]:
    """Ragged batch in one C call (used for the timed CPU baseline)."""
    lib = _load()
    flat = np.concatenate([np.ascontiguousarray(c, dtype=np.float32).ravel() for c in costs]) if costs else np.zeros(0, np.float32)
    nr = np.array([c.shape[0] for c in costs], dtype=np.int64)
    nc = np.array([c.shape[1] for c in costs], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(nr * nc)[:-1]]).astype(np.int64) if len(costs) else np.zeros(0, np.int64)
    n_out = np.minimum(nr, nc)
    out_off = np.concatenate([[0], np.cumsum(n_out)[:-1]]).astype(np.int64) if len(costs) else np.zeros(0, np.int64)
    rows = np.empty(int(n_out.sum()), dtype=np.int64)
    cols = np.empty_like(rows)
    i64p = ctypes.POINTER(ctypes.c_int64)
    rc = lib.lsap_oracle_batch_f32(flat.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), off.ctypes.data_as(i64p),
                                   nr.ctypes.data_as(i64p), nc.ctypes.data_as(i64p), len(costs),
                                   out_off.ctypes.data_as(i64p), rows.ctypes.data_as(i64p), cols.ctypes.data_as(i64p))
    _raise(rc)
    return [(rows[o:o + k].copy(), cols[o:o + k].copy()) for o, k in zip(out_off, n_out)]


def linear_sum_assignment_py(cost) -> tuple[np.ndarray, np.ndarray]:
    """Pure-Python twin of oracle/lsap.c (small cases only; documents the tie-break rule)."""
    c = np.asarray(cost, dtype=np.float64)
    nr, nc = c.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    flip = nc < nr
    if flip:
        c = c.T.copy()
        nr, nc = nc, nr
    if np.isnan(c).any() or np.isneginf(c).any():
        raise ValueError("matrix contains invalid numeric entries")
    u = [0.0] * nr
    v = [0.0] * nc
    col_of_row = [-1] * nr
    row_of_col = [-1] * nc
    pred = [-1] * nc
    for cur in range(nr):
        todo = [nc - 1 - t for t in range(nc)]
        dist = [float("inf")] * nc
        rows_seen, cols_seen = set(), set()
        reach, i, sink = 0.0, cur, -1
        while sink < 0:
            rows_seen.add(i)
            best, best_t = float("inf"), -1
            for t, j in enumerate(todo):
                r = ((reach + float(c[i, j])) - u[i]) - v[j]
                if r < dist[j]:
                    dist[j], pred[j] = r, i
                if dist[j] < best or (dist[j] == best and row_of_col[j] < 0):
                    best, best_t = dist[j], t
            reach = best
            if reach == float("inf"):
                raise ValueError("cost matrix is infeasible")
            j = todo[best_t]
            if row_of_col[j] < 0:
                sink = j
            else:
                i = row_of_col[j]
            cols_seen.add(j)
            todo[best_t] = todo[-1]
            todo.pop()
        u[cur] += reach
        for a in rows_seen:
            if a != cur:
                u[a] += reach - dist[col_of_row[a]]
        for b in cols_seen:
            v[b] -= reach - dist[b]
        j = sink
        while True:
            a = pred[j]
            row_of_col[j] = a
            col_of_row[a], j = j, col_of_row[a]
            if a == cur:
                break
    if flip:
        pairs = [(b, row_of_col[b]) for b in range(nc) if row_of_col[b] >= 0]
    else:
        pairs = [(a, col_of_row[a]) for a in range(nr)]
    return (np.array([p[0] for p in pairs], dtype=np.int64), np.array([p[1] for p in pairs], dtype=np.int64))
