"""Philox4x32-10 counter-based RNG -- numpy restatement of the stream contract.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, smoke() and
bench.py's cpu_baseline leg.  The CUDA library implements the same contract in
csrc/philox.cuh; the two must agree bit for bit so that Pauli trajectories can be
compared trajectory by trajectory, not only statistically.

Contract (shared with the GPU library):
    key      = (seed & 0xffffffff, seed >> 32)
    counter  = (index, stream, traj & 0xffffffff, traj >> 32)
    stream   = 0: noise site `index` of trajectory `traj`
               1: measurement sample `index` of trajectory `traj`
    uniform  = ((x0 | x1 << 32) >> 11) * 2**-53      (x0, x1 = first two output words)

The algorithm is the published Philox4x32 with 10 rounds (Salmon et al., SC'11); the
reference itself draws its randomness inside qiskit-aer (third party, not in tree), so
only the *distribution* is pinned by the reference, never the stream.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)

STREAM_NOISE = 0
STREAM_MEASURE = 1
STREAM_READOUT = 2      # readout-error flip of classical bit `index` of shot `traj`


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def uniform(seed, index, stream, traj):
    """53-bit uniform doubles in [0,1) for (index, stream, traj) under `seed` (all broadcastable)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    traj = np.asarray(traj, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(
        np.asarray(index, dtype=np.uint64).astype(np.uint32),
        np.asarray(stream, dtype=np.uint32),
        (traj & _MASK).astype(np.uint32),
        (traj >> np.uint64(32)).astype(np.uint32),
        seed & 0xFFFFFFFF,
        seed >> 32,
    )
    bits = x0.astype(np.uint64) | (x1.astype(np.uint64) << np.uint64(32))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def pauli_from_uniform(u, px, py, pz):
    """Map uniforms to Pauli codes 0=I, 1=X, 2=Y, 3=Z with cumulative order X, Y, Z, (rest) I."""
    u = np.asarray(u)
    code = np.zeros(u.shape, dtype=np.uint8)
    code[u < px + py + pz] = 3
    code[u < px + py] = 2
    code[u < px] = 1
    return code
