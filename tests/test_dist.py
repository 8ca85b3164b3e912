"""N>1 path on CPU: world_size-2 gloo.  Each rank processes its shard (here with the oracle
standing in for the kernels -- the host-side sharding and base-scan logic is what is tested),
exchanges 16 bytes, and the merged uint32 index space must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.pyref import OracleLib
from vg_renderer_b200 import workloads as W
from vg_renderer_b200.dist import exclusive_bases, shard_range
from vg_renderer_b200.tables import BatchTables

N = 600


def _merged(out, vertex_base):
    idx = out["index"].astype(np.uint64)
    for m in range(out["num_meshes"]):
        i0, i1 = int(out["mesh_idx_off"][m]), int(out["mesh_idx_off"][m + 1])
        idx[i0:i1] += out["mesh_vtx_off"][m] + np.uint64(vertex_base)
    return idx


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(N, rank, world)
    out = OracleLib().run_batch(BatchTables.cubic_strokes(W.config3_cubics(hi - lo, first=lo), width=4.0, cap=1, join=1))
    vb, ib, tv, ti = exclusive_bases(out["num_vertices"], out["num_indices"])
    q.put((rank, lo, hi, vb, ib, tv, ti, _merged(out, vb), out["pos"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merged_index_space_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    whole = OracleLib().run_batch(BatchTables.cubic_strokes(W.config3_cubics(N), width=4.0, cap=1, join=1))
    assert res[0][1:3] == (0, N // 2) and res[1][1:3] == (N // 2, N)
    assert res[0][3] == 0 and res[0][4] == 0
    assert res[1][3] == len(res[0][8]) and res[1][4] == len(res[0][7])
    assert res[0][5] == res[1][5] == whole["num_vertices"] and res[0][6] == res[1][6] == whole["num_indices"]
    np.testing.assert_array_equal(np.concatenate([res[0][7], res[1][7]]), _merged(whole, 0))
    np.testing.assert_array_equal(np.concatenate([res[0][8], res[1][8]]), whole["pos"])


def test_shard_range_covers_everything():
    for n, w in ((10, 3), (1_000_000, 8), (5, 8), (0, 2)):
        r = [shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert torch.__version__


# ---- the same on GPUs, through the library's own group (NCCL all-gather inside vgb_batch_run_merged) -------
def _gpu_worker(rank, world, uid_path, q):
    import time
    import vg_renderer_b200 as vgb
    if rank == 0:
        uid = vgb.Group.unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.05)
        uid = open(uid_path, "rb").read()
    g = vgb.Group(rank, rank, world, uid)
    lo, hi = shard_range(N, rank, world)
    b = vgb.Batch(rank)
    b.add_cubic_strokes(W.config3_cubics(hi - lo, first=lo), width=4.0, cap=1, join=1)
    b.run_merged(g)
    info = b.merged_info()
    out = b.download()
    b.run_merged(g)  # a re-run (sizes known: fully asynchronous) must give the same
    out2 = b.download()
    assert np.array_equal(out["index"], out2["index"])
    res = g.gather(b, owner=0)
    merged = None
    if rank == 0:
        V, I = res["total_vertices"], res["total_indices"]
        merged = (vgb.copy_from_device(res["pos"], V * 8, np.float32).reshape(-1, 2), vgb.copy_from_device(res["color"], V * 4, np.uint32),
                  vgb.copy_from_device(res["index"], I * 4, np.uint32))
    q.put((rank, info, out["index"], merged, res["ms"]))
    time.sleep(0.5)
    g.close()


@pytest.mark.gpu
def test_two_gpu_merged_buffer_equals_single_process(tmp_path):
    """vgb_batch_run_merged on 2 ranks (NCCL all-gather of the (V, I) totals inside the library, base added on the
    device) + vgb_group_gather (payload into rank 0's buffer over NVLink) == the single-process uint32 result."""
    import vg_renderer_b200 as vgb
    if vgb.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    uid_path = str(tmp_path / "uid")
    procs = [ctx.Process(target=_gpu_worker, args=(r, 2, uid_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    b = vgb.Batch(0)
    b.add_cubic_strokes(W.config3_cubics(N), width=4.0, cap=1, join=1)
    b.run(index_u32=True)
    whole = b.download()
    assert res[0][1]["vertex_base"] == 0 and res[1][1]["vertex_base"] == res[0][1]["total_vertices"] - (whole["num_vertices"] - res[1][1]["vertex_base"])
    assert res[0][1]["total_vertices"] == whole["num_vertices"] and res[1][1]["total_indices"] == whole["num_indices"]
    np.testing.assert_array_equal(np.concatenate([res[0][2], res[1][2]]), whole["index"])
    pos, col, idx = res[0][3]
    np.testing.assert_array_equal(idx, whole["index"])
    np.testing.assert_array_equal(pos, whole["pos"])
    np.testing.assert_array_equal(col, whole["color"])
