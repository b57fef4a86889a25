"""Philox4x32-10 uniforms on the host (numpy), same counter contract as csrc/dtc_hd.cuh:philox_uniform.

Used for classical post-processing that never touches a state: readout-error bit flips (stream 2: index = classical bit,
trajectory word = shot).  Streams 0 (Pauli sites) and 1 (measurement samples) are drawn on the device."""
import numpy as np

STREAM_READOUT = 2
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_LO = np.uint64(0xFFFFFFFF)


def uniform(seed, index, stream, traj):
    """[0,1) doubles with 53 random bits for counter (index, stream, traj lo, traj hi) and key = seed (broadcasting)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    traj = np.asarray(traj, dtype=np.uint64)
    c = [np.asarray(index, dtype=np.uint64) & _LO, np.asarray(stream, dtype=np.uint64) & _LO, traj & _LO, traj >> np.uint64(32)]
    c = [np.ascontiguousarray(x) for x in np.broadcast_arrays(*c)]
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]                    # 32 x 32 -> 64 bit products (operands < 2^32)
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & _LO, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & _LO]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    bits = c[0] | (c[1] << np.uint64(32))
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
