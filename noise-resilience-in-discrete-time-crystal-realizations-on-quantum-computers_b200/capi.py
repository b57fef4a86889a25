"""ctypes binding of the C ABI in include/dtcsim.h (libdtcsim.so, built in-tree by __graft_entry__.build()).

There is no CPU fallback: if the CUDA library is missing or a call fails, a RuntimeError/ValueError is
raised with the library's own message (the analogue of the exceptions Aer raises from run(), fast.py:211).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DTCSIM_LIB") or os.path.join(_HERE, "csrc", "libdtcsim.so")   # env override: tuning builds

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64 = ctypes.c_int64
c_u64 = ctypes.c_uint64
c_f64p = ctypes.POINTER(ctypes.c_double)
c_vp = ctypes.c_void_p

_SIGS = {
    "dtc_version": (ctypes.c_int, []),
    "dtc_last_error": (ctypes.c_char_p, []),
    "dtc_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "dtc_program_destroy": (ctypes.c_int, [c_vp]),
    "dtc_program_set_events": (ctypes.c_int, [c_vp, c_i64, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_f64p, c_f64p,
                                              ctypes.c_double]),
    "dtc_program_set_exec_layers": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "dtc_program_set_readout_hint": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "dtc_program_finalize": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dtc_program_num_passes": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_workspace_bytes": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(ctypes.c_size_t)]),
    "dtc_program_run": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_u64, c_u64, c_u64, c_vp, ctypes.c_size_t, c_vp]),
    "dtc_program_prepare": (ctypes.c_int, [c_vp, c_i64, c_i64, c_u64, c_vp, ctypes.c_size_t, c_vp]),
    "dtc_program_run_passes": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i64, c_u64, c_u64,
                                              c_vp, ctypes.c_size_t, c_vp]),
    "dtc_program_pass_info": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_frames": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                          ctypes.POINTER(c_vp)]),
    "dtc_program_set_readout": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, c_i32p, ctypes.c_int,
                                               c_i32p, ctypes.c_int, c_i32p]),
    "dtc_program_readout": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "dtc_program_set_fused_rdm": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_fused_rdm": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_vp)]),
    "dtc_set_stream_engine": (ctypes.c_int, [ctypes.c_int]),
    "dtc_set_high_stride_bit": (ctypes.c_int, [ctypes.c_int]),
    "dtc_set_stream_ctas": (ctypes.c_int, [ctypes.c_int]),
    "dtc_set_resident_bytes": (ctypes.c_int, [ctypes.c_size_t]),
    "dtc_program_resident_info": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                                 ctypes.POINTER(ctypes.c_size_t)]),
    "dtc_program_run_resident": (ctypes.c_int, [c_vp, c_vp, ctypes.c_size_t, c_i64, c_i64, c_u64, c_u64, c_u64, c_vp,
                                                ctypes.c_size_t, c_vp]),
    "dtc_program_last_run_info": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_num_stream_passes": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_last_run_flags": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_set_profiling": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "dtc_program_pass_time": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "dtc_program_pass_times": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_int)]),
    "dtc_materialize": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "dtc_probs": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_i32p, c_vp, c_vp, c_vp]),
    "dtc_rdm": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_i32p, c_vp, c_vp]),
    "dtc_expect_z": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, c_vp, c_vp]),
    "dtc_sample_rows": (ctypes.c_int, [c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_u64, c_i64, c_vp, c_vp]),
    "dtc_sample_states": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_u64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "dtc_sample_states_multi": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_u64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "dtc_dm_init": (ctypes.c_int, [c_vp, ctypes.c_int, c_u64, c_vp]),
    "dtc_dm_rot": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_vp]),
    "dtc_dm_diag": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_i32p, c_f64p, ctypes.c_int, c_i32p, c_i32p,
                                   c_f64p, c_vp]),
    "dtc_dm_pauli_channel": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, c_vp]),
    "dtc_dm_superop": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_f64p, c_vp]),
    "dtc_dm_probs": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_i32p, c_vp, c_vp]),
    "dtc_dm_run": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_i32p, c_i32p, c_i32p, c_i32p, c_f64p, c_f64p,
                                  ctypes.POINTER(ctypes.c_int), c_vp]),
    "dtc_shard_pack": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_i32p, c_vp]),
    "dtc_shard_unpack": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_i32p, c_vp]),
}
EXPORTED = tuple(_SIGS)

ENGINE_AUTO, ENGINE_GENERIC, ENGINE_TILE = 0, 1, 2
INIT_KEEP, INIT_ZERO = 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFE
_lib = None


def load():
    """Load libdtcsim.so; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"dtcsim CUDA library not found at {LIB_PATH}; build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def set_high_stride_bit(bit=15):
    """Groups starting at or above this internal bit use five-qubit tiles of 2 KB runs (default 15)."""
    check(load().dtc_set_high_stride_bit(int(bit)))


def set_stream_ctas(n=0):
    """Persistent CTAs per k_tile_stream launch (0: one per SM)."""
    check(load().dtc_set_stream_ctas(int(n)))


def set_resident_bytes(nbytes=64 << 20):
    """State bytes a resident run keeps in flight (trajectories per group x state size; L2 is 126 MB)."""
    check(load().dtc_set_resident_bytes(int(nbytes)))


# Module switch: evolve() uses resident execution (k_tile_resident: every sweep of a circuit in one persistent launch over
# groups of trajectories whose states stay in L2) for read-out-only runs of eligible programs.  Off by default: on B200 it is
# 15-20 % slower than one launch per sweep on config C2 (profiles/experiments_r2.md) -- it draws 800 W instead of 990 W and
# keeps the maximum SM clock, but a group small enough for L2 (4 states of 16 MiB) leaves too little distance between a tile
# and the tiles it depends on.  Environment DTCSIM_RESIDENT=1 turns it on.
RESIDENT = os.environ.get("DTCSIM_RESIDENT", "0") == "1"


def set_stream_engine(enable):
    """True/False: force k_tile_stream on/off for eligible passes; None: library default (on)."""
    check(load().dtc_set_stream_engine(-1 if enable is None else int(bool(enable))))


def check(rc):
    if rc != 0:
        msg = load().dtc_last_error().decode(errors="replace")
        if rc in (-1, -3):
            raise ValueError(f"dtcsim: {msg}")
        raise RuntimeError(f"dtcsim (status {rc}): {msg}")


def i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


class ProgramHandle:
    """Owns a dtc_program*; created from a plan.Program."""

    def __init__(self, prog, device, engine=ENGINE_AUTO, n_local=None):
        lib = load()
        self.prog = prog
        self.n_local = getattr(prog, "n_main", prog.n) if n_local is None else n_local
        self.device = device
        self._h = c_vp()
        check(lib.dtc_program_create(prog.n, prog.n_layers, ctypes.byref(self._h)))
        try:
            ev = prog.arrays()
            keep = [i32(ev["type"]), i32(ev["layer"]), i32(ev["q0"]), i32(ev["q1"]), i32(ev["slot"]),
                    f64(ev["val"]), f64(ev["probs"])]
            check(lib.dtc_program_set_events(self._h, len(ev["type"]), *[k[1] for k in keep],
                                             float(prog.global_phase)))
            check(lib.dtc_program_set_exec_layers(self._h, int(prog.n_exec_layers)))
            small = getattr(prog, "small", None)
            if small is not None and len(small["reg_bits"]) == 1:
                check(lib.dtc_program_set_readout_hint(self._h, int(small["reg_bits"][0])))
            check(lib.dtc_program_finalize(self._h, int(device), int(engine), int(self.n_local)))
            small = getattr(prog, "small", None)
            if small is not None:
                idx = np.ascontiguousarray(small["events"], dtype=np.int64)
                rb, rbp = i32(small["reg_bits"] if small["reg_bits"] else [0])
                eb, ebp = i32(small["elim_bits"])
                mb, mbp = i32([b for b, _c in prog.measures])
                check(lib.dtc_program_set_readout(self._h, len(idx), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                  len(small["reg_bits"]), rbp, len(small["elim_bits"]), ebp,
                                                  len(prog.measures), mbp))
        except Exception:
            self.close()
            raise

    @property
    def num_passes(self):
        n = ctypes.c_int(0)
        check(load().dtc_program_num_passes(self._h, ctypes.byref(n)))
        return n.value

    @property
    def num_stream_passes(self):
        """Passes of the schedule that run on the TMA-fed k_tile_stream (the rest use k_tile_pass)."""
        n = ctypes.c_int(0)
        check(load().dtc_program_num_stream_passes(self._h, ctypes.byref(n)))
        return n.value

    def workspace_bytes(self, n_traj):
        b = ctypes.c_size_t(0)
        check(load().dtc_program_workspace_bytes(self._h, int(n_traj), ctypes.byref(b)))
        return b.value

    def run(self, state_ptr, n_traj, traj_offset, seed, ws_ptr, ws_bytes, stream, init_index=0, rank_bits=0):
        check(load().dtc_program_run(self._h, state_ptr, int(n_traj), int(traj_offset),
                                     int(seed) & 0xFFFFFFFFFFFFFFFF, int(init_index), int(rank_bits),
                                     ws_ptr, ws_bytes, stream))

    def resident_info(self, n_traj):
        """(eligible, trajectories per group, scratch bytes) for dtc_program_run_resident."""
        e, g, b = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_size_t(0)
        check(load().dtc_program_resident_info(self._h, int(n_traj), ctypes.byref(e), ctypes.byref(g), ctypes.byref(b)))
        return bool(e.value), g.value, b.value

    def run_resident(self, scratch_ptr, scratch_bytes, n_traj, traj_offset, seed, ws_ptr, ws_bytes, stream, init_index=0,
                     rank_bits=0):
        check(load().dtc_program_run_resident(self._h, scratch_ptr, int(scratch_bytes), int(n_traj), int(traj_offset),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, int(init_index), int(rank_bits),
                                              ws_ptr, ws_bytes, stream))

    def last_run_info(self):
        """(resident execution used, kernels launched) of the last run."""
        r, k = ctypes.c_int(0), ctypes.c_int(0)
        check(load().dtc_program_last_run_info(self._h, ctypes.byref(r), ctypes.byref(k)))
        return bool(r.value), k.value

    def prepare(self, n_traj, traj_offset, seed, ws_ptr, ws_bytes, stream):
        """Frame walk only (sign masks into the workspace); run_passes() then executes pass ranges."""
        check(load().dtc_program_prepare(self._h, int(n_traj), int(traj_offset), int(seed) & 0xFFFFFFFFFFFFFFFF, ws_ptr,
                                         ws_bytes, stream))

    def run_passes(self, state_ptr, begin, end, n_traj, ws_ptr, ws_bytes, stream, store_last=None, n_ctas=0,
                   init_index=INIT_KEEP, rank_bits=0):
        check(load().dtc_program_run_passes(self._h, state_ptr, store_last, int(begin), int(end), int(n_ctas), int(n_traj),
                                            int(init_index), int(rank_bits), ws_ptr, ws_bytes, stream))

    def pass_info(self, i):
        """(runs on the streaming engine, tiles contiguous in memory) of pass i."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        check(load().dtc_program_pass_info(self._h, int(i), ctypes.byref(a), ctypes.byref(b)))
        return bool(a.value), bool(b.value)

    def set_fused_rdm(self, enable=True):
        """Let the last pass reduce the read-out qubit's density matrix instead of storing the state (factorised
        circuits whose last pass runs on k_tile_stream).  Returns True if the next run() will do so."""
        active = ctypes.c_int(0)
        check(load().dtc_program_set_fused_rdm(self._h, int(bool(enable)), ctypes.byref(active)))
        return bool(active.value)

    def fused_rdm_ptr(self, ws_ptr, n_traj):
        ptr = c_vp()
        check(load().dtc_program_fused_rdm(self._h, ws_ptr, int(n_traj), ctypes.byref(ptr)))
        return ptr.value

    def readout(self, rdm_ptr, ws_ptr, n_traj, probs_ptr, stream):
        check(load().dtc_program_readout(self._h, rdm_ptr, ws_ptr, int(n_traj), probs_ptr, stream))

    def set_profiling(self, enable=True):
        check(load().dtc_program_set_profiling(self._h, int(bool(enable))))

    def pass_time(self):
        """(milliseconds, launches) of the state-sweep loop of the last run (waits for it)."""
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        check(load().dtc_program_pass_time(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def pass_times(self):
        """[(milliseconds, tile layout mode)] per pass of the last whole-program run (profiling on; waits for it)."""
        cap = 4096
        ms = (ctypes.c_float * cap)()
        modes = (ctypes.c_int * cap)()
        n = ctypes.c_int(0)
        check(load().dtc_program_pass_times(self._h, ms, modes, cap, ctypes.byref(n)))
        return [(ms[i], modes[i]) for i in range(min(n.value, cap))]

    def last_run_flags(self):
        """(first pass generated the state, last pass fused the read-out) of the last run()."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        check(load().dtc_program_last_run_flags(self._h, ctypes.byref(a), ctypes.byref(b)))
        return bool(a.value), bool(b.value)

    def frames(self, ws_ptr, n_traj):
        fx, fz, ph = c_vp(), c_vp(), c_vp()
        check(load().dtc_program_frames(self._h, ws_ptr, int(n_traj), ctypes.byref(fx), ctypes.byref(fz),
                                        ctypes.byref(ph)))
        return fx.value, fz.value, ph.value

    def close(self):
        if self._h:
            load().dtc_program_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
