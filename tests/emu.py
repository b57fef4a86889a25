"""ctypes wrapper of the CPU emulation harness tests/emul/emul.cpp (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cpp")
OUT = os.path.join(HERE, "emul", "_build", "libdtcemu.so")
CSRC = os.path.join(os.path.dirname(HERE),
                    "noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("dtc_hd.cuh", "dtc_core.hpp", "dtc_stream.cuh", "dtc_readout.cuh", "dtc_dm.cuh")]
    if force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, SRC])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def run(prog, n_traj=1, traj_offset=0, seed=0, init_index=0, engine=0, n_local=None, rank_bits=0):
    """Execute a compiled Program on the emulator. Returns (psi' [T,2^n_local], fx, fz, ph, n_passes)."""
    ev = prog.arrays()
    n_local = prog.n_main if n_local is None else n_local
    state = np.zeros((n_traj, 1 << n_local), dtype=np.complex128)
    fx = np.zeros(n_traj, dtype=np.uint64)
    fz = np.zeros(n_traj, dtype=np.uint64)
    ph = np.zeros(n_traj, dtype=np.int32)
    npass = ctypes.c_int(0)
    err = ctypes.create_string_buffer(512)
    rc = lib().emu_run(
        ctypes.c_int(prog.n), ctypes.c_int(prog.n_layers), ctypes.c_int64(len(ev["type"])),
        _p(ev["type"], ctypes.c_int32), _p(ev["layer"], ctypes.c_int32), _p(ev["q0"], ctypes.c_int32),
        _p(ev["q1"], ctypes.c_int32), _p(ev["slot"], ctypes.c_int32), _p(ev["val"], ctypes.c_double),
        _p(ev["probs"], ctypes.c_double), ctypes.c_double(prog.global_phase), ctypes.c_int(engine),
        ctypes.c_int(n_local), ctypes.c_int(prog.n_exec_layers), ctypes.c_int64(n_traj), ctypes.c_int64(traj_offset), ctypes.c_uint64(seed),
        ctypes.c_uint64(init_index), ctypes.c_uint64(rank_bits),
        state.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), _p(fx, ctypes.c_uint64), _p(fz, ctypes.c_uint64),
        _p(ph, ctypes.c_int32), ctypes.byref(npass), err, ctypes.c_int(512))
    if rc != 0:
        raise ValueError(err.value.decode())
    return state, fx, fz, ph, npass.value


def schedule(prog, n_local=None, cap=4096):
    ev = prog.arrays()
    n_local = prog.n_main if n_local is None else n_local
    rows = np.zeros((cap, 22), dtype=np.int32)
    err = ctypes.create_string_buffer(512)
    n = lib().emu_schedule(
        ctypes.c_int(prog.n), ctypes.c_int(prog.n_layers), ctypes.c_int64(len(ev["type"])),
        _p(ev["type"], ctypes.c_int32), _p(ev["layer"], ctypes.c_int32), _p(ev["q0"], ctypes.c_int32),
        _p(ev["q1"], ctypes.c_int32), _p(ev["slot"], ctypes.c_int32), _p(ev["val"], ctypes.c_double),
        _p(ev["probs"], ctypes.c_double), ctypes.c_int(n_local), ctypes.c_int(prog.n_exec_layers), _p(rows, ctypes.c_int32), ctypes.c_int(cap),
        err, ctypes.c_int(512))
    if n < 0:
        raise ValueError(err.value.decode())
    return rows[:min(n, cap)], n


def readout_small(prog, rdm, masks, fx):
    """csrc/dtc_readout.cuh on the CPU: rdm [T,2^k,2^k], masks [n_layers,4,T] uint64 (all layers), fx [T] -> probs [T,2^m]."""
    ev = prog.arrays()
    sm = prog.small
    T = rdm.shape[0]
    m = len(prog.measures)
    idx = np.ascontiguousarray(sm["events"], dtype=np.int64)
    rb = np.ascontiguousarray(sm["reg_bits"] if sm["reg_bits"] else [0], dtype=np.int32)
    eb = np.ascontiguousarray(sm["elim_bits"], dtype=np.int32)
    mb = np.ascontiguousarray([b for b, _c in prog.measures], dtype=np.int32)
    rdm = np.ascontiguousarray(rdm, dtype=np.complex128)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    assert masks.shape == (prog.n_layers, 4, T)
    fx = np.ascontiguousarray(fx, dtype=np.uint64)
    out = np.zeros((T, 1 << m))
    err = ctypes.create_string_buffer(512)
    rc = lib().emu_readout_small(
        ctypes.c_int(prog.n), ctypes.c_int(prog.n_layers), ctypes.c_int64(len(ev["type"])),
        _p(ev["type"], ctypes.c_int32), _p(ev["layer"], ctypes.c_int32), _p(ev["q0"], ctypes.c_int32),
        _p(ev["q1"], ctypes.c_int32), _p(ev["slot"], ctypes.c_int32), _p(ev["val"], ctypes.c_double),
        _p(ev["probs"], ctypes.c_double), ctypes.c_int64(len(idx)), _p(idx, ctypes.c_int64),
        ctypes.c_int(len(sm["reg_bits"])), _p(rb, ctypes.c_int32), ctypes.c_int(len(eb)), _p(eb, ctypes.c_int32),
        ctypes.c_int(m), _p(mb, ctypes.c_int32), rdm.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        _p(masks, ctypes.c_uint64), _p(fx, ctypes.c_uint64), ctypes.c_int64(T),
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), err, ctypes.c_int(512))
    if rc != 0:
        raise ValueError(err.value.decode())
    return out


def set_high_stride_bit(bit=15):
    lib().emu_set_high_stride_bit(ctypes.c_int(bit))


def dm_run(n, flat_segments, rho0=None, reg_passes=True, wide13=True):
    """dtc_dm_run on the CPU (planner + per-thread code of csrc/dtc_dm.cuh).  flat_segments = backend.flatten_dm_segments(...).
    Returns (rho [2^n cols, 2^n rows], info dict)."""
    st, so, q0, q1, val, pr = flat_segments
    rho = np.zeros(1 << (2 * n), dtype=np.complex128)
    if rho0 is None:
        rho[0] = 1.0
    else:
        rho[:] = np.asarray(rho0, dtype=np.complex128).reshape(-1)
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    info = np.zeros(4, dtype=np.int32)
    err = ctypes.create_string_buffer(512)
    rc = lib().emu_dm_run(
        ctypes.c_int(n), ctypes.c_int(len(st)), _p(st, ctypes.c_int32), _p(so, ctypes.c_int32), _p(q0, ctypes.c_int32),
        _p(q1, ctypes.c_int32), _p(val, ctypes.c_double), pr.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        ctypes.c_int(int(reg_passes)), ctypes.c_int(int(wide13)), rho.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        _p(info, ctypes.c_int32), err, ctypes.c_int(512))
    if rc != 0:
        raise ValueError(err.value.decode())
    return rho.reshape(1 << n, 1 << n), {"sweeps": int(info[0]), "reg_passes": int(info[1]), "worst_conflict": int(info[2])}


def dm_program(prog, reg_passes=True, wide13=True):
    """backend.run_density_matrix on the CPU: stretches of R / D / N segments through the emulated dtc_dm_run, non-Pauli
    channels through the emulated k_dm_superop.  Returns rho [2^n cols, 2^n rows]."""
    from dtcsim.backend import flatten_dm_segments, split_dm_segments
    n = prog.n
    rho = np.zeros(1 << (2 * n), dtype=np.complex128)
    rho[0] = 1.0
    for kind, payload in split_dm_segments(prog.dm_segments):
        if kind == "K":
            for q, S in payload:
                Sf = np.ascontiguousarray(np.asarray(S, dtype=np.complex128).reshape(16)).view(np.float64)
                lib().emu_dm_superop(ctypes.c_int(n), ctypes.c_int(int(q)), _p(Sf, ctypes.c_double),
                                     rho.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        else:
            rho = dm_run(n, flatten_dm_segments(payload), rho0=rho, reg_passes=reg_passes, wide13=wide13)[0].reshape(-1).copy()
    return rho.reshape(1 << n, 1 << n)
