"""Prototype (CPU, uses the oracle) for the next flatten design: one thread per COMMAND instead of per sub-path.

Claim to check: the only thing command c needs from its predecessors is the path's last kept vertex E_c
(P1 of a curve, the dedupe partner of its first vertex).  For moveTo / lineTo / cubicTo / quadTo the previous
command's nominal end point N_{c-1} is known from the command data alone, and E_c == N_{c-1} unless the
epsilon-dedupe (path.cpp:761-784) dropped that end point.  So: flatten every command on its own from
E_c := N_{c-1}, keep a flag "my last candidate vertex was kept" per command, and the concatenation equals the
sequential walk whenever all flags of a sub-path are set; otherwise (and for arc / arcTo, whose end point is
computed) that sub-path falls back to the sequential walk.

    python tools/proto_command_parallel_flatten.py
prints, per scene, how many sub-paths are eligible and whether every eligible one reproduces the sequential
result bit for bit."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle.pyref import OracleLib  # noqa: E402
from vg_renderer_b200 import workloads as W  # noqa: E402
from vg_renderer_b200.tables import (CMD_CLOSE, CMD_CUBIC_TO, CMD_LINE_TO, CMD_MOVE_TO, CMD_QUAD_TO)  # noqa: E402

ARITY = {0: 2, 1: 2, 2: 6, 3: 4, 4: 5, 5: 4, 6: 5, 7: 8, 8: 3, 9: 4, 10: 6, 12: 0}
SIMPLE = {CMD_MOVE_TO, CMD_LINE_TO, CMD_CUBIC_TO, CMD_QUAD_TO, CMD_CLOSE}


def split_subpaths(cmds, data):
    """(cmd list, data list) per sub-path of ONE path: a new sub-path at every moveTo (the recorder's rule for
    the simple commands)"""
    out, cur, k = [], [], 0
    for c in cmds:
        n = ARITY[int(c)]
        if int(c) == CMD_MOVE_TO and cur:
            out.append(cur)
            cur = []
        cur.append((int(c), [float(v) for v in data[k:k + n]]))
        k += n
    if cur:
        out.append(cur)
    return out


def check_scene(name, t, orc):
    elig = total = equal = 0
    for p in range(t.num_paths):
        c0, c1 = int(t.path_cmd_off[p]), int(t.path_cmd_off[p + 1])
        d0, d1 = int(t.path_data_off[p]), int(t.path_data_off[p + 1])
        cmds, data = t.cmd[c0:c1], t.data[d0:d1]
        if any(int(c) == 11 for c in cmds):  # polyline command: variable arity, not in these scenes
            continue
        scale, tol = float(t.path_scale[p]), float(t.path_tol[p])
        for sub in split_subpaths(cmds, data):
            total += 1
            if any(c not in SIMPLE for c, _ in sub) or sub[0][0] != CMD_MOVE_TO:
                continue
            # sequential result (the oracle walks the sub-path as the reference does)
            seq, _ = orc.flatten([c for c, _ in sub], [v for _, d in sub for v in d], scale, tol)
            closed = sub[-1][0] == CMD_CLOSE
            body = sub[:-1] if closed else sub
            # per-command, independent: flatten [moveTo E_c, command c]; E_c := nominal end of command c-1
            pieces, ok = [np.asarray([body[0][1]], np.float32)], True
            end = body[0][1]
            for c, d in body[1:]:
                pts, _ = orc.flatten([CMD_MOVE_TO, c], list(end) + d, scale, tol)
                new = pts[1:]
                nominal = np.asarray(d[-2:], np.float32)
                if len(new) == 0 or not np.array_equal(new[-1], nominal):
                    ok = False  # the end point was deduped away: E of the next command is not N_c
                    break
                pieces.append(new)
                end = d[-2:]
            if not ok:
                continue
            elig += 1
            cat = np.concatenate(pieces)
            if closed:  # pathClose drops a last vertex that coincides with the first (path.cpp:707-726)
                dx, dy = np.float32(cat[-1, 0] - cat[0, 0]), np.float32(cat[-1, 1] - cat[0, 1])
                if len(cat) > 2 and np.float32(dx * dx + dy * dy) < np.float32(1e-5):
                    cat = cat[:-1]
            equal += int(cat.shape == seq.shape and np.array_equal(cat, seq))
    print(f"{name:28s} sub-paths {total:6d}  eligible {elig:6d} ({100.0 * elig / max(total, 1):5.1f} %)  "
          f"bit-identical to the sequential walk {equal:6d} / {elig}")
    return elig == equal


if __name__ == "__main__":
    from cases import fuzz_recorder  # noqa: E402
    o = OracleLib()
    good = check_scene("tiger-like", W.tiger_like(), o)
    good &= check_scene("tiger-like x6 instances", W.instance_tables(W.tiger_like(), 6), o)
    for seed in (1, 2, 3):
        good &= check_scene(f"fuzz seed {seed}", fuzz_recorder(seed, 200).tables(), o)
    print("OK" if good else "MISMATCH")
