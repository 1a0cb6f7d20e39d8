"""CPU oracle -- counter-based random streams (TEST INFRASTRUCTURE).

The reference is unseeded: spawn points come from the global ``np.random.uniform``
(exp02_vFinal_task.py:583-607), hit rolls from the global ``random.random``
(gun.py:94) and motor noise from a shared ``np.random.RandomState``
(quadcopter.py:137,181).  "Identical seeds" parity therefore treats every draw
as data: both the oracle and the CUDA kernels read Philox4x32-10 streams keyed
by (seed, env, stream, index), so a draw is a pure function of its coordinates.

  counter = (index, stream | sub << 8, env, 0)      key = (seed_lo, seed_hi)

Streams: HIT (one uniform per shot actually fired, index = per-env running
count), SPAWN (one uniform per angle drawn, index = per-env running count),
MOTOR (four normals per drone per physics substep; sub = drone slot, index =
per-env physics-substep count), FUSE (threatsense LiDAR fusion draws).
"""
from __future__ import annotations

import numpy as np

STREAM_HIT, STREAM_SPAWN, STREAM_MOTOR, STREAM_FUSE = 1, 2, 3, 4

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10; all arguments broadcastable uint32 arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + _W0); k1 = np.uint32(k1 + _W1)
    return c0, c1, c2, c3


def u01(x):
    """uint32 -> uniform in [0,1) on a 24-bit grid (exact in fp32 and fp64)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def uniform(seed: int, env, stream: int, index, sub=0):
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c1 = np.asarray(sub, dtype=np.uint32) * np.uint32(256) + np.uint32(stream)
    x0, _, _, _ = philox4x32(index, c1, env, 0, k0, k1)
    return u01(x0)


def normal4(seed: int, env, index, sub):
    """Four standard normals (Box-Muller on the four Philox words)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c1 = np.asarray(sub, dtype=np.uint32) * np.uint32(256) + np.uint32(STREAM_MOTOR)
    x = philox4x32(index, c1, env, 0, k0, k1)
    out = []
    for a, b in ((x[0], x[1]), (x[2], x[3])):
        u1 = ((a >> np.uint32(8)).astype(np.float64) + 1.0) * (1.0 / 16777216.0)
        u2 = u01(b)
        r = np.sqrt(-2.0 * np.log(u1))
        out += [r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)]
    return np.stack(out, axis=-1)
