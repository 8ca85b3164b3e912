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


# ---- configs 2 and 5: a tiger-LIKE scene -------------------------------------------------------
# The Ghostscript tiger.svg is not in the reference and cannot be fetched here (no network), so
# BASELINE configs[1]/[4] are run on a deterministic synthetic stand-in with the same structure:
# ~240 paths, each 1-3 closed sub-paths of cubic/line segments, every path convex-AA-filled and
# most of them AA-stroked with a thin SVG-style width (butt caps, miter joins).
def tiger_like(num_paths=240, seed=7):
    """-> tables.BatchTables of one scene instance (identity transform, scale 1)"""
    from .tables import Recorder
    rng = np.random.default_rng(seed)
    r = Recorder(fringe=1.0)
    for p in range(num_paths):
        r.begin_path(1.0, 0.25)
        cx0, cy0 = rng.uniform(50, 950, 2)
        for s in range(int(rng.integers(1, 4))):
            cx, cy = cx0 + rng.uniform(-30, 30), cy0 + rng.uniform(-30, 30)
            k = int(rng.integers(3, 41))
            rad = rng.uniform(4, 120)
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            rr = rad * rng.uniform(0.6, 1.4, k)
            px, py = cx + rr * np.cos(ang), cy + rr * np.sin(ang)
            r.move_to(px[0], py[0])
            for i in range(1, k + 1):
                j = i % k
                if rng.random() < 0.7:
                    h = rng.uniform(0.2, 0.5) * rad * 0.5
                    r.cubic_to(px[i - 1] + rng.uniform(-h, h), py[i - 1] + rng.uniform(-h, h),
                               px[j] + rng.uniform(-h, h), py[j] + rng.uniform(-h, h), px[j], py[j])
                else:
                    r.line_to(px[j], py[j])
            r.close()
        col = 0xFF000000 | int(rng.integers(0, 1 << 24))
        r.convex_fill(col, True)
        if rng.random() < 0.6:
            r.stroke(0, 0xFF000000 | int(rng.integers(0, 1 << 24)), float(rng.uniform(1.05, 3.0)), 0, 0)
    return r.tables()


def instance_tables(t, num_instances, seed=11, fringe=1.0, first=0, total=None):
    """config 5: the scene x num_instances, each with T(U[0,4096)^2) * R(U[0,2pi)) * S(U[0.5,2)).
    first / total: this is the slice [first, first + num_instances) of a stream of `total` instances (ranks shard by
    range and must draw the same transform for the same instance whatever the world size).
    Reference semantics (vg.cpp:2969-2981, 3416-3433): the path is flattened with scale = avgScale,
    stroke width = clamp(width * avgScale, 0, 200), thin stroke when that is <= fringe."""
    from .tables import BatchTables
    rng = np.random.default_rng(seed)
    P, M = t.num_paths, num_instances
    out = BatchTables(fringe)
    out.num_paths = P * M
    ncmd, ndat, ndraw = int(t.path_cmd_off[-1]), int(t.path_data_off[-1]), int(t.draw_begin[-1])
    k = np.arange(M, dtype=np.uint64)[:, None]
    out.path_cmd_off = np.concatenate([(t.path_cmd_off[:-1][None, :] + k * ncmd).reshape(-1), [M * ncmd]]).astype(np.uint32)
    out.path_data_off = np.concatenate([(t.path_data_off[:-1][None, :] + k * ndat).reshape(-1), [M * ndat]]).astype(np.uint32)
    out.draw_begin = np.concatenate([(t.draw_begin[:-1][None, :] + k * ndraw).reshape(-1), [M * ndraw]]).astype(np.uint32)
    out.cmd = np.tile(t.cmd, M)
    out.data = np.tile(t.data, M)
    T = M + first if total is None else total
    sl = slice(first, first + M)
    s = rng.uniform(0.5, 2.0, T).astype(np.float32)[sl]
    a = rng.uniform(0, 2 * np.pi, T)[sl]
    tx, ty = rng.uniform(0, 4096, T)[sl], rng.uniform(0, 4096, T)[sl]
    m = np.stack([s * np.cos(a), s * np.sin(a), -s * np.sin(a), s * np.cos(a), tx, ty], 1).astype(np.float32)
    out.path_xform = np.repeat(m, P, axis=0).reshape(-1)
    out.path_scale = np.repeat(s, P)
    out.path_tol = np.tile(t.path_tol, M)
    w = np.clip(np.tile(t.draw_width, M) * np.repeat(s, ndraw), 0.0, 200.0).astype(np.float32)
    op = np.tile(t.draw_op, M)
    thin = (op == 0) & (w <= fringe)
    op = np.where(thin, 3, op).astype(np.uint32)
    out.draw_op = op
    out.draw_width = np.where(thin, np.float32(fringe), w).astype(np.float32)
    out.draw_color = np.tile(t.draw_color, M)
    out.draw_cap = np.tile(t.draw_cap, M)
    out.draw_join = np.tile(t.draw_join, M)
    return out
