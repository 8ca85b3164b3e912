"""Synthetic inputs for BASELINE.json's configs. All deterministic, numpy only.

RNG: xorshift64 (s^=s<<13; s^=s>>7; s^=s<<17), float = ((s>>40)&0xFFFFFF)/2^24, one independent
stream per item seeded with splitmix64(seed + item index) so a batch can be generated vectorised
and any sub-range reproduces exactly (ranks shard by item range)."""
import numpy as np

SEED = 0x9E3779B97F4A7C15
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return z ^ (z >> np.uint64(31))


def _xorshift_floats(first, count, per_item, seed=SEED):
    """[count, per_item] floats in [0,1): item i uses its own xorshift64 stream."""
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + count, dtype=np.uint64)
        s = _splitmix(idx + np.uint64(seed))
        s[s == 0] = np.uint64(1)
        out = np.empty((count, per_item), np.float32)
        for k in range(per_item):
            s ^= (s << np.uint64(13)) & _M
            s ^= s >> np.uint64(7)
            s ^= (s << np.uint64(17)) & _M
            out[:, k] = ((s >> np.uint64(40)) & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(16777216.0)
    return out


def config1_cubic():
    """single cubic M(100,100) C(200,0)(300,400)(400,100)"""
    return np.array([[100, 100, 200, 0, 300, 400, 400, 100]], np.float32)


def config3_cubics(n, first=0, seed=SEED):
    """n random cubics: p0 ~ U[0,1024)^2, p1..p3 = p0 + U[-128,128)^2 -> [n, 8] float32"""
    u = _xorshift_floats(first, n, 8, seed)
    p = np.empty((n, 8), np.float32)
    p[:, 0:2] = u[:, 0:2] * np.float32(1024.0)
    for k in range(1, 4):
        p[:, 2 * k:2 * k + 2] = p[:, 0:2] + (u[:, 2 * k:2 * k + 2] * np.float32(256.0) - np.float32(128.0))
    return p


def config4_rrects(n, r, first=0):
    """n rounded rects pathRoundedRect(10 + i%7, 10, 4r+50, 3r+20, r) -> [n, 5] float32"""
    i = np.arange(first, first + n)
    q = np.empty((n, 5), np.float32)
    q[:, 0] = 10 + (i % 7)
    q[:, 1] = 10
    q[:, 2] = 4 * r + 50
    q[:, 3] = 3 * r + 20
    q[:, 4] = r
    return q
