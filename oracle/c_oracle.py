"""ctypes wrapper of oracle/dtc_oracle.c (TEST INFRASTRUCTURE; see oracle/oracle.py for provenance)."""
import ctypes
import os
import subprocess

import numpy as np

from . import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dtc_oracle.c")
OUT = os.path.join(HERE, "_build", "liboracle_c.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(SRC) > os.path.getmtime(OUT):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"])
    return OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_prob1.restype = ctypes.c_double
        _lib.orc_threads.restype = ctypes.c_int
    return _lib


def threads():
    return int(lib().orc_threads())


def set_threads(n=None):
    """Use n OpenMP threads (default: every host core this process may run on).  torchrun exports OMP_NUM_THREADS=1 to
    its ranks; the CPU baseline of bench.py runs on rank 0 alone and should use the whole host."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_threads(ctypes.c_int(int(n)))
    return threads()


def encode(ops):
    """Compacted op tuples -> flat arrays for orc_run."""
    ops = [O._norm_op(o) for o in ops]
    n = len(ops)
    kind = np.full(n, 2, dtype=np.int32)
    q0 = np.zeros(n, dtype=np.int32)
    q1 = np.zeros(n, dtype=np.int32)
    mats = np.zeros((n, 8), dtype=np.float64)
    for i, (name, qs, params, _) in enumerate(ops):
        if name in O.ONE_QUBIT:
            kind[i] = 0
            q0[i] = qs[0]
            mats[i] = O.gate_matrix(name, params).reshape(4).view(np.float64)
        elif name == "cx":
            kind[i] = 1
            q0[i], q1[i] = qs
        elif name in ("measure", "barrier"):
            q0[i] = qs[0] if qs else 0
        else:
            raise ValueError(f"c_oracle: lower {name} to u/cx first")
    return kind, q0, q1, mats


def run_state(ops, n, pauli_codes=None, init=0, out=None):
    """Gate-by-gate evolution of one state; pauli_codes[i] is applied after op i. Returns psi (2^n)."""
    kind, q0, q1, mats = encode(ops)
    psi = np.empty(1 << n, dtype=np.complex128) if out is None else out
    L = lib()
    L.orc_init(psi.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int64(init))
    pc = None
    if pauli_codes is not None:
        pc = np.ascontiguousarray(pauli_codes, dtype=np.uint8)
    L.orc_run(psi.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int64(len(kind)),
              kind.ctypes.data_as(ctypes.c_void_p), q0.ctypes.data_as(ctypes.c_void_p),
              q1.ctypes.data_as(ctypes.c_void_p), mats.ctypes.data_as(ctypes.c_void_p),
              pc.ctypes.data_as(ctypes.c_void_p) if pc is not None else None)
    return psi


def run_trajectory(ops, n, noise, seed, traj, out=None):
    """One noisy trajectory with the Philox contract of oracle.sample_paulis."""
    ops = [O._norm_op(o) for o in ops]
    sites, codes = O.sample_paulis(ops, noise, seed, [traj])
    per_op = np.zeros(len(ops), dtype=np.uint8)
    for s, (i, _, _) in enumerate(sites):
        per_op[i] = codes[0, s]
    return run_state(ops, n, per_op, out=out)


def prob1(psi, n, q):
    return float(lib().orc_prob1(psi.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), ctypes.c_int(q)))
