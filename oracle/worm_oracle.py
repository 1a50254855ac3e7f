"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libworm_oracle.so (worm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libworm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "worm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.worm_oracle_mt.restype = C.c_int
        _lib.worm_oracle_philox_run.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def mt_stream(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().worm_oracle_mt_stream(C.c_uint32(seed & 0xFFFFFFFF), C.c_int64(n), _p(out, C.c_uint32))
    return out


def bond_map(L: int) -> np.ndarray:
    rows = np.zeros((4 * L * L, 5), dtype=np.int64)
    lib().worm_oracle_bond_map(C.c_int(L), _p(rows, C.c_int64))
    return rows


def mt_tables(T: float):
    inc = np.zeros(256, dtype=np.uint64)
    dec = np.zeros(256, dtype=np.uint64)
    lib().worm_oracle_mt_tables(C.c_double(T), _p(inc, C.c_uint64), _p(dec, C.c_uint64))
    return inc, dec


def thresholds(T: float):
    k, t, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().worm_oracle_thresholds(C.c_double(T), C.byref(k), C.byref(t), C.byref(h))
    return k.value, t.value, h.value


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().worm_oracle_philox(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def run_mt(L: int, T: float, num_steps: int, seed: float, bond_flag: int = 1,
           max_events: int = 4096, max_trace: int = 0) -> dict:
    """Deterministic mode, one chain (restates worm_ising_2d.cpp:111-157)."""
    nb2 = 2 * L * L
    out = np.zeros(4, dtype=np.int64)
    final = np.zeros(nb2, dtype=np.int32)
    n_events = C.c_int64(0)
    n_trace = C.c_int64(0)
    events = np.zeros((max(max_events, 1), 3), dtype=np.int64)
    dumps = np.zeros((max(max_events, 1), nb2), dtype=np.int32) if bond_flag else None
    trace = np.zeros((max(max_trace, 1), 2), dtype=np.int64)
    rc = lib().worm_oracle_mt(C.c_int(L), C.c_double(T), C.c_uint64(int(num_steps)), C.c_double(seed),
                              C.c_int(bond_flag), _p(out, C.c_int64), _p(final, C.c_int32),
                              C.c_int64(max_events), C.byref(n_events), _p(events, C.c_int64),
                              _p(dumps, C.c_int32), C.c_int64(max_trace), C.byref(n_trace),
                              _p(trace, C.c_int64))
    if rc != 0:
        raise RuntimeError("worm_oracle_mt rc=%d" % rc)
    ne = min(n_events.value, max_events)
    return dict(Z=int(out[0]), Nb_tot=int(out[1]), step_num=int(out[2]), Nb_final=int(out[3]),
                final_bonds=final, n_events=n_events.value, events=events[:ne].copy(),
                dumps=(dumps[:ne].copy() if dumps is not None else None),
                n_trace=n_trace.value, trace=trace[:min(n_trace.value, max_trace)].copy())


def run_philox(L: int, algo: int, chain_id: int, seed: int, T: float, step_begin: int, n_steps: int,
               therm: int, bonds: np.ndarray | None = None, head: int = 0, tail: int = 0,
               acc: np.ndarray | None = None, hist: bool = False, dump_every: int = 0,
               max_dumps: int = 0, thr=None) -> dict:
    """Production mode, one chain (CPU restatement of the CUDA path; see worm_oracle.c)."""
    N = L * L
    kfix, tfix, hfix = thr if thr is not None else thresholds(T)
    b = np.zeros(2 * N, dtype=np.uint8) if bonds is None else np.ascontiguousarray(bonds, dtype=np.uint8).copy()
    h, t = C.c_int32(head), C.c_int32(tail)
    a = np.zeros(8, dtype=np.int64) if acc is None else np.ascontiguousarray(acc, dtype=np.int64).copy()
    hs = np.zeros(N, dtype=np.int64) if hist else None
    nd = C.c_int64(0)
    ev = np.zeros((max(max_dumps, 1), 3), dtype=np.int64)
    dm = np.zeros((max(max_dumps, 1), 2 * N), dtype=np.uint8) if max_dumps else None
    rc = lib().worm_oracle_philox_run(C.c_int(L), C.c_int(algo), C.c_uint64(chain_id), C.c_uint64(seed),
                                      C.c_uint64(kfix), C.c_uint64(tfix), C.c_uint64(hfix),
                                      C.c_uint64(step_begin), C.c_uint64(n_steps), C.c_uint64(therm),
                                      _p(b, C.c_uint8), C.byref(h), C.byref(t), _p(a, C.c_int64),
                                      _p(hs, C.c_int64), C.c_uint64(dump_every), C.c_int64(max_dumps),
                                      C.byref(nd), _p(ev, C.c_int64), _p(dm, C.c_uint8))
    if rc != 0:
        raise RuntimeError("worm_oracle_philox_run rc=%d" % rc)
    k = min(nd.value, max_dumps)
    return dict(bonds=b, head=h.value, tail=t.value, acc=a, hist=hs, n_dumps=nd.value,
                events=ev[:k].copy(), dumps=(dm[:k].copy() if dm is not None else None))
