#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

  metric   Mverts/s emitted (AA stroke+fill), whole-job aggregate over all ranks
  --workload (config.workload names it; the default is the configuration the metric is quoted on)
    cubics        configs[2]: 1M random cubic beziers per GPU (p0~U[0,1024)^2, p1..p3 = p0 +
                  U[-128,128)^2), batched flatten + AA stroke, width 4, round caps + round joins, per-mesh uint16
                  indices.  Rank r takes items [r*1M, (r+1)*1M) of the stream (weak scaling).
    tiger         configs[1]: a tiger-LIKE scene (240 paths: fills + AA strokes; the Ghostscript tiger.svg is neither in
                  the reference nor fetchable here -- vg_renderer_b200/workloads.py:tiger_like), x --instances (default
                  1) with random affine transforms, every rank the same scene range (weak scaling).
    rrect         configs[3]: 1M rounded-rect convex AA fills per GPU, sweep over the corner radius r in
                  {2,8,32,128,512,2048} (12..205 polygon points); one step = the six batches back to back.
    tiger_x10000  configs[4]: the tiger-like scene x 10000 instances (random T*R*S per instance, re-tessellated at each
                  instance's scale as vg.cpp does), SHARDED over the ranks (strong scaling: total work fixed), ONE
                  merged uint32 vertex/index space: every step runs vgb_batch_run_merged, i.e. the NCCL all-gather of
                  the ranks' (V, I) totals sits between the count stage and the emit kernel.
  --merged  (cubics / tiger) the same merged space for the weak-scaling workloads.
  --gather  (with a merged space, N > 1) additionally place every rank's payload in ONE buffer on rank 0 over NVLink
            (vgb_group_gather), once, after the timed steps; reported as "nvlink" (GB/s into the owner), not part of value.

  step     one pass of the hot path over the batch(es) (vgb_batch_run: flatten -> count -> scan -> emit)
  value    inputs already resident in HBM; device time by CUDA events on the batch's own stream
           (event before step 0, event after step K-1: inter-step gaps and the run's internal
           size read-backs are included), max over ranks
  e2e      same metric through the C ABI with HOST buffers: record (host tables / the caller's pinned coordinate
           array) + H2D + run + D2H of pos/colour/index/mesh table into pinned host arrays, every step, wall clock
           between barriers, max over ranks; pcie_frac = its D2H rate / a plain pinned D2H copy measured in the same run
  e2e_device_consumer   the same call sequence for a consumer that lives ON the device (final-layout streams read
           through vgb_batch_device_ptrs; no D2H of the mesh): host tables in, device pointers out
  roofline the dominant kernel of the workload: algorithmic bytes (input + 12*V + s*I) / its CUDA-event
           duration, against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline / --impl reference: the reference's own path.cpp + stroker.cpp (oracle/_ref,
           compiled from /root/reference against the bx shim) on the host cores

Every workload writes far more than the 126 MB L2 per step (stated in config.l2), except the single tiger scene,
for which a 256 MB buffer is overwritten between steps.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CUBICS_PER_GPU = 1_000_000
WIDTH, CAP, JOIN = 4.0, 1, 1  # round caps + round joins
METRIC = "Mverts/s emitted (AA stroke+fill)"
UNIT = "Mverts/s"
RRECT_RADII = (2, 8, 32, 128, 512, 2048)


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


# ------------------------------------------------------------------------------------------
# clocks (NVML sampled in a thread during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.ok = False
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref): one process per host core, disjoint item ranges
# ------------------------------------------------------------------------------------------
_BARRIER = None


def _cpu_init(barrier):
    global _BARRIER
    _BARRIER = barrier


def _ref_tables(workload, first, count):
    """the reference-side description of `count` items of the workload's stream starting at item `first`"""
    from vg_renderer_b200 import workloads as W
    from vg_renderer_b200.tables import BatchTables
    if workload == "cubics":
        return BatchTables.cubic_strokes(W.config3_cubics(count, first=first), width=WIDTH, cap=CAP, join=JOIN)
    if workload == "rrect":
        # items cycle through the six radii so that any range carries the sweep's mix
        per = max(count // len(RRECT_RADII), 1)
        parts = [BatchTables.rounded_rect_fills(W.config4_rrects(per, r, first=first)) for r in RRECT_RADII]
        return parts
    # tiger-like scene x count instances (seeded by the range so that ranges differ)
    scene = W.tiger_like()
    return scene if (workload == "tiger" and count == 1) else W.instance_tables(scene, count, seed=11 + first)


def _cpu_worker(args):
    kind, workload, first, count = args
    from oracle.pyref import RefLib
    lib = RefLib(kind)
    t = _ref_tables(workload, first, count)
    parts = t if isinstance(t, list) else [t]
    for p in parts:
        lib.run_batch(p, store=False, path_range=(0, min(p.num_paths, 2000)))  # warm caches / allocator
    _BARRIER.wait()  # every process starts the timed part together
    t0 = time.perf_counter()
    verts = 0
    for p in parts:
        verts += lib.run_batch(p, store=False)["num_vertices"]
    t1 = time.perf_counter()
    return verts, t0, t1


def cpu_reference_run(workload, per_proc, procs, kind):
    """-> (total_vertices, wall_seconds, per-process seconds): one process per core on disjoint ranges of the stream,
    all released by a barrier; wall = first start to last finish (CLOCK_MONOTONIC is system-wide)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(barrier,)) as pool:
        res = pool.map(_cpu_worker, [(kind, workload, i * per_proc, per_proc) for i in range(procs)], chunksize=1)
    verts = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    each = sorted(r[2] - r[1] for r in res)
    return verts, wall, each


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref_kind():
    from oracle.pyref import have_ref
    # stroke code is identical in the SIMD=0/1 builds; use the stock (SIMD=1) configuration
    return "sse" if have_ref("sse") else ("scalar" if have_ref("scalar") else None)


# items of the workload's stream one reference process works through per step (bounded: the arm must end in minutes)
REF_PER_PROC = {"cubics": 50_000, "rrect": 60_000, "tiger": 20, "tiger_x10000": 20}


def spread(each):
    """per-process wall seconds of the reference arm: min / median / max (oversubscription shows as a wide spread)"""
    return {"min_s": round(each[0], 3), "median_s": round(each[len(each) // 2], 3), "max_s": round(each[-1], 3)}


def workload_text(args, n):
    if args.workload == "cubics":
        return (f"{n} random cubic beziers per GPU (SURVEY 8d config 3), batched flatten + AA stroke, width 4, round caps + round joins, "
                f"{'merged uint32' if args.merged else 'per-mesh uint16'} indices")
    if args.workload == "rrect":
        return f"{n} rounded-rect convex AA fills per GPU and radius, radius sweep r in {list(RRECT_RADII)} (SURVEY 8d config 4), per-mesh uint16 indices"
    if args.workload == "tiger":
        return (f"tiger-like scene (240 paths, convex AA fills + AA strokes; synthetic stand-in for tiger.svg) x {n} instance(s) per GPU, "
                f"{'merged uint32' if args.merged else 'per-mesh uint16'} indices (SURVEY 8d config 2)")
    return f"tiger-like scene x {n} instances in total, random affine per instance, sharded over the ranks, one merged uint32 index space (SURVEY 8d config 5)"


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    kind = ref_kind()
    cores = host_cores()
    wl = args.workload
    per_proc = REF_PER_PROC[wl]
    if wl == "cubics" and args.ref_curves:
        per_proc = max(2000, min(100_000, args.ref_curves // max(cores, 1)))
    if kind is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/*.so missing (make -C oracle ref needs /root/reference)"}))
        return 0
    times, verts, each_all = [], 0, []
    for s in range(args.warmup + args.steps):
        v, w, each = cpu_reference_run(wl, per_proc, cores, kind)
        if s >= args.warmup:
            times.append(w)
            verts += v
            each_all = each
    total = sum(times)
    value = verts / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "strong" if wl == "tiger_x10000" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args, default_items(args)),
                   "reference_sample": f"{cores} processes x {per_proc} items of the same stream per step", "process_wall": spread(each_all)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{cores} procs x {per_proc} items/step, oracle/_ref/libvgref{'_sse' if kind == 'sse' else ''}.so "
                                   f"(reference path.cpp+stroker.cpp, SIMD={'1' if kind == 'sse' else '0'})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def default_items(args):
    if args.items:
        return args.items
    return {"cubics": args.curves, "rrect": 1_000_000, "tiger": args.instances or 1, "tiger_x10000": args.instances or 10_000}[args.workload]


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """One process per GPU: run this rank on the CPUs next to its GPU, so that the pinned host buffers of
    the end-to-end leg (allocated later, first-touched by this process) live on the GPU's own NUMA node and
    the ranks' D2H streams do not all cross the socket interconnect.  Host-side placement only."""
    info = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        info = {"cpus": len(cpus), "first_cpu": cpus[0] if cpus else None}
        try:
            info["gpu_numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:  # noqa: BLE001
            pass
        try:
            node = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and
                    os.path.exists(f"/sys/devices/system/node/{d}/cpu{cpus[0]}")]
            info["cpu_numa_node"] = int(node[0][4:]) if node else None
        except Exception:  # noqa: BLE001
            pass
    except Exception as e:  # no NVML / no permission: keep the default placement
        info = {"error": str(e)[:80]}
    return info


class GpuWorkload:
    """the resident batches of one rank + how a step runs them"""

    def __init__(self, args, rank, world, local):
        import vg_renderer_b200 as vgb
        from vg_renderer_b200 import workloads as W
        self.vgb, self.args = vgb, args
        self.batches, self.labels, self.tables, self.input_bytes = [], [], [], []
        self.merged = args.merged or args.workload == "tiger_x10000"
        wl = args.workload
        n = default_items(args)
        self.items_text = n
        if wl == "cubics":
            self.p8 = W.config3_cubics(n, first=rank * n)
            b = vgb.Batch(local)
            b.reset(1.0)
            b.add_cubic_strokes(self.p8, width=WIDTH, cap=CAP, join=JOIN)
            self.batches.append(b)
            self.labels.append("cubics")
            self.input_bytes.append(32 * n)
        elif wl == "rrect":
            for r in RRECT_RADII:
                b = vgb.Batch(local)
                b.reset(1.0)
                b.add_rounded_rect_fills(W.config4_rrects(n, r, first=rank * n))
                self.batches.append(b)
                self.labels.append(f"r={r}")
                self.input_bytes.append(20 * n)
        else:
            scene = W.tiger_like()
            if wl == "tiger":
                t = scene if n == 1 else W.instance_tables(scene, n, seed=11 + rank)
            else:
                from vg_renderer_b200.dist import shard_range
                lo, hi = shard_range(n, rank, world)
                # every instance draws its transform from the same stream whatever the sharding: generate all, slice
                t = W.instance_tables(scene, hi - lo, seed=11, first=lo, total=n)
                self.shard = (lo, hi)
            b = vgb.Batch(local)
            b.add_tables(t)
            self.batches.append(b)
            self.labels.append(wl)
            self.tables.append(t)
            self.input_bytes.append(int(t.cmd.nbytes + t.data.nbytes + (t.path_xform.nbytes if t.path_xform is not None else 0) + 8 * t.num_paths))
        for b in self.batches:
            b.upload()
        self.group = None

    def step(self):
        for b in self.batches:
            if self.group is not None:
                b.run_merged(self.group)
            else:
                b.run(index_u32=self.merged)

    def sync(self):
        for b in self.batches:
            b.sync()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vgb", choices=["vgb", "reference"])
    ap.add_argument("--workload", default="cubics", choices=["cubics", "tiger", "rrect", "tiger_x10000"])
    ap.add_argument("--curves", type=int, default=CUBICS_PER_GPU, help="cubics per GPU (default: BASELINE config)")
    ap.add_argument("--items", type=int, default=0, help="items per GPU of the chosen workload (rects per radius / instances)")
    ap.add_argument("--instances", type=int, default=0, help="tiger: instances per GPU (default 1); tiger_x10000: instances in total (default 10000)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-curves", type=int, default=0, help="reference arm: total cubics per step (0 = 50k per core)")
    ap.add_argument("--merged", action="store_true", help="single merged uint32 index space: in-library all-gather of per-rank (V,I) totals")
    ap.add_argument("--no-merged-leg", action="store_true", help="N>1: skip the extra merged-space measurement of a plain run")
    ap.add_argument("--gather", action="store_true", help="merged space, N>1: also gather the payload into rank 0's buffer over NVLink (once, reported separately)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vgb" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    numa = bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of timings + the 128-byte group id
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import vg_renderer_b200 as vgb
    from vg_renderer_b200.api import pinned_empty

    if vgb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")

    def barrier():
        if dist is not None:
            dist.barrier()

    def allred(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(x):
        return allred(x, dist.ReduceOp.MAX) if dist is not None else x

    def allsum(x):
        return allred(x, dist.ReduceOp.SUM) if dist is not None else x

    if dist is not None:
        every = [None] * world
        dist.all_gather_object(every, numa)  # one entry per rank: where each rank's pinned buffers will live
        numa = every
    wl = GpuWorkload(args, rank, world, local)
    merged = wl.merged
    if merged and world > 1:
        # the library's own communicator: rank 0 makes the id, torch.distributed only carries the 128 bytes
        import torch
        ids = [vgb.Group.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        wl.group = vgb.Group(local, rank, world, ids[0])

    for _ in range(args.warmup):
        wl.step()
    wl.sync()
    flush = None
    if args.workload == "tiger" and default_items(args) < 64:
        flush = vgb.Batch(local)  # a second batch whose output (>= 256 MB) evicts the scene's lines between steps
        from vg_renderer_b200 import workloads as W
        flush.add_cubic_strokes(W.config3_cubics(70_000), width=WIDTH, cap=CAP, join=JOIN)
        flush.run()
        flush.sync()
    b0 = wl.batches[0]
    sampler = ClockSampler(local)
    barrier()
    wl.sync()
    sampler.start()
    w0 = time.perf_counter()
    dev_ms = 0.0
    if flush is None and len(wl.batches) == 1:
        b0.mark(0)
        for _ in range(args.steps):
            wl.step()
        b0.mark(1)
        dev_ms = b0.elapsed_ms(0, 1)
    elif flush is None:
        # several batches (own streams): run them one after the other, device time = sum of the batches' event spans
        for b in wl.batches:
            b.mark(0)
            for _ in range(args.steps):
                b.run(index_u32=merged)
            b.mark(1)
            b.sync()
        dev_ms = sum(b.elapsed_ms(0, 1) for b in wl.batches)
    else:
        for _ in range(args.steps):
            flush.run()
            flush.sync()
            b0.mark(0)
            wl.step()
            b0.mark(1)
            dev_ms += b0.elapsed_ms(0, 1)
    wl.sync()
    w1 = time.perf_counter()
    clocks = sampler.stop()
    barrier()
    sizes = [b.sizes() for b in wl.batches]
    stage = [b.last_run_ms() for b in wl.batches]
    # per-kernel duration of the dominant kernel: a few extra (untimed) steps, CUDA events around emit
    emit_ms = [[] for _ in wl.batches]
    for _ in range(5):
        for k, b in enumerate(wl.batches):
            if wl.group is not None:
                b.run_merged(wl.group)
            else:
                b.run(index_u32=merged)
            emit_ms[k].append(b.last_run_ms()["emit"])
    emit_avg = [float(np.mean(e)) for e in emit_ms]

    idx_bytes = 4 if merged else 2
    V = [int(s.num_vertices) for s in sizes]
    I = [int(s.num_indices) for s in sizes]
    alg = [wl.input_bytes[k] + 12 * V[k] + idx_bytes * I[k] for k in range(len(V))]
    ms_per_step = allmax(dev_ms) / args.steps
    wall_ms_per_step = allmax((w1 - w0) * 1e3) / args.steps
    total_verts = allsum(float(sum(V)))
    value = total_verts / (ms_per_step * 1e-3) / 1e6

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = sum(alg) / (sum(emit_avg) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "emit_traffic.json")
    if os.path.exists(tpath) and args.workload == "cubics":
        try:
            tj = json.load(open(tpath))
            if tj.get("curves") == default_items(args) and tj.get("index_bytes") == idx_bytes:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
    kernel = "k_fill_emit" if args.workload == "rrect" else "k_group_emit"
    tot_stage = {k: float(sum(s[k] for s in stage)) for k in stage[0]}
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": sum(alg), "kernel_ms": sum(emit_avg),
                "pipeline_frac": (sum(alg) / (ms_per_step * 1e-3) / 1e9) / peak,
                "stage_ms": {"flatten": tot_stage["flatten"], "count_scan": tot_stage["stroke_count"], "emit": tot_stage["emit"], "total": tot_stage["total"]}}
    if len(wl.batches) > 1:
        roofline["per_batch"] = [{"batch": wl.labels[k], "verts": V[k], "indices": I[k], "ms": stage[k]["total"], "emit_ms": emit_avg[k],
                                  "mverts_s": V[k] / stage[k]["total"] / 1e3, "emit_gbs": alg[k] / emit_avg[k] / 1e6, "emit_frac": alg[k] / emit_avg[k] / 1e6 / peak}
                                 for k in range(len(V))]

    # ---- NVLink: payload gather into one owner (merged space only) --------------------------
    nvlink = None
    minfo = None
    if wl.group is not None:
        minfo = b0.merged_info()
        if args.gather:
            barrier()
            cold = wl.group.gather(b0, owner=0)  # first use: NCCL sets up its point-to-point channels, the owner allocates
            barrier()
            g = wl.group.gather(b0, owner=0)
            moved = (g["total_vertices"] - V[0]) * 12 + (g["total_indices"] - I[0]) * 4 if rank == 0 else 0
            ms = allmax(g["ms"])
            moved = allmax(float(moved))
            nvlink = {"what": "vgb_group_gather: every rank's pos/colour/index into rank 0's merged buffer (ncclSend/ncclRecv)", "ms": ms, "first_call_ms": allmax(cold["ms"]),
                      "bytes_into_owner": moved, "gbs_into_owner": moved / (ms * 1e-3) / 1e9, "reference_gbs": 770.0,
                      "frac_of_measured_peer_copy": moved / (ms * 1e-3) / 1e9 / 770.0}

    # ---- e2e: host buffers in, host buffers out, every step --------------------------------
    e2e = None
    e2e_dev = None
    if not args.no_e2e:
        e2e, e2e_dev = run_e2e(args, wl, vgb, pinned_empty, V, I, idx_bytes, merged, barrier, allmax, total_verts)

    # ---- N > 1, plain run: also measure the ONE exchange step of the path -----------------------------------
    # (so that a default multi-GPU run carries the data-path collective: merged uint32 space through the library group)
    merged_space = None
    if world > 1 and wl.group is None and args.workload == "cubics" and not args.no_merged_leg:
        import torch  # noqa: F401
        ids = [vgb.Group.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        grp = vgb.Group(local, rank, world, ids[0])
        for _ in range(3):
            b0.run_merged(grp)
        b0.sync()
        barrier()
        k = 20
        b0.mark(2)
        for _ in range(k):
            b0.run_merged(grp)
        b0.mark(3)
        mms = allmax(b0.elapsed_ms(2, 3)) / k
        mi = b0.merged_info()
        barrier()
        grp.gather(b0, owner=0)
        barrier()
        g = grp.gather(b0, owner=0)
        gms = allmax(g["ms"])
        moved = allmax(float((g["total_vertices"] - V[0]) * 12 + (g["total_indices"] - I[0]) * 4 if rank == 0 else 0))
        merged_space = {"what": "the same batches as ONE merged uint32 vertex/index space: vgb_batch_run_merged = count stage -> ncclAllGather of 16 B/rank "
                                "(V,I totals) on the batch stream -> device-side base -> emit", "value": total_verts / (mms * 1e-3) / 1e6, "unit": UNIT,
                        "ms_per_step": mms, "steps": k, "nranks": world, "total_vertices": mi["total_vertices"], "total_indices": mi["total_indices"],
                        "gather_into_rank0": {"what": "vgb_group_gather (ncclSend/ncclRecv over NVLink)", "ms": gms, "bytes": moved,
                                              "gbs": moved / (gms * 1e-3) / 1e9, "frac_of_measured_peer_copy_770": moved / (gms * 1e-3) / 1e9 / 770.0}}
        grp.close()

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        kind = ref_kind()
        if kind is not None:
            cores = host_cores()
            per = REF_PER_PROC[args.workload] * (2 if args.workload == "cubics" else 1)
            v, wsec, each = cpu_reference_run(args.workload, per, cores, kind)
            cpu = {"value": v / wsec / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{cores} procs x {per} items of the same stream, oracle/_ref/libvgref{'_sse' if kind == 'sse' else ''}.so, "
                             f"pathReset..strokerXXX only, {wsec:.2f} s wall", "process_wall": spread(each)}
            v1, w1s, _ = cpu_reference_run(args.workload, per, 1, kind)
            cpu["one_core_value"] = v1 / w1s / 1e6

    if rank == 0:
        launches = int(sum(s.kernel_launches for s in sizes))
        coll = ""
        if merged and world > 1:
            coll = " + ncclAllGather of 16 B/rank (V,I totals) per step inside vgb_batch_run_merged -> device-side vertex base"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.workload == "tiger_x10000" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(args, default_items(args)),
                       "verts_per_gpu": sum(V), "indices_per_gpu": sum(I), "points_per_gpu": int(sum(s.num_points for s in sizes)),
                       "meshes_per_gpu": int(sum(s.num_meshes for s in sizes)),
                       "l2": ("a 280 MB mesh is written between steps (L2 flush)" if flush is not None else
                              "each step writes ~%.2f GB per GPU (>> 126 MB L2): no explicit flush" % (sum(alg) / 1e9)),
                       "parallelism": (f"{world} x sharded instance ranges" if args.workload == "tiger_x10000" else f"{world} x independent path ranges")
                                      + (coll if coll else ", no data-path collective" + ("; the one exchange step (merged uint32 space: ncclAllGather of the "
                                         "(V,I) totals inside vgb_batch_run_merged, payload gather over NVLink) is measured in merged_space" if merged_space else "")),
                       "merged": minfo, "host_placement": numa},
            "clocks": clocks, "e2e": e2e, "e2e_device_consumer": e2e_dev, "gpu_launches": launches * args.steps,
            "launches_per_step": launches, "wall_ms_per_step": wall_ms_per_step,
            "roofline": roofline, "nvlink": nvlink, "merged_space": merged_space, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        if wl.group is not None:
            wl.group.close()
        dist.destroy_process_group()
    return 0


def run_e2e(args, wl, vgb, pinned_empty, V, I, idx_bytes, merged, barrier, allmax, total_verts):
    """host buffers in -> host buffers out through the C ABI, every step"""
    nb = len(wl.batches)
    Vm, Im = max(V), max(I)
    pos, _p1 = pinned_empty((Vm + 16, 2), np.float32)
    col, _p2 = pinned_empty((Vm + 16,), np.uint32)
    idx, _p3 = pinned_empty((Im + 16,), np.uint32 if merged else np.uint16)
    nmesh = max(int(b.sizes().num_meshes) for b in wl.batches)
    mvo, _p5 = pinned_empty((nmesh + 1,), np.uint64)
    mio, _p6 = pinned_empty((nmesh + 1,), np.uint64)
    pin_in = None
    if args.workload == "cubics":
        pin_in, _p4 = pinned_empty(wl.p8.shape, np.float32)
        pin_in[:] = wl.p8
    from vg_renderer_b200 import workloads as W

    def rerecord(k, b):
        b.reset(1.0)
        if args.workload == "cubics":
            b.set_cubic_strokes(pin_in, width=WIDTH, cap=CAP, join=JOIN)  # borrowed, DMA'd by upload
        elif args.workload == "rrect":
            b.add_rounded_rect_fills(wl.rr[k])
        else:
            b.add_tables(wl.tables[k])
        b.upload()

    if args.workload == "rrect":
        n = default_items(args)
        rk = env_int("RANK", 0)
        wl.rr = [W.config4_rrects(n, r, first=rk * n) for r in RRECT_RADII]

    def run(b):
        if wl.group is not None:
            b.run_merged(wl.group)
        else:
            b.run(index_u32=merged)

    def e2e_step():
        for k, b in enumerate(wl.batches):
            rerecord(k, b)
            run(b)
            b._chk(b.L.vgb_batch_download(b.h, pos.ctypes.data, col.ctypes.data, idx.ctypes.data, len(pos), len(idx),
                                          mvo.ctypes.data, mio.ctypes.data))

    def dev_step():
        for k, b in enumerate(wl.batches):
            rerecord(k, b)
            run(b)
            b.sync()
            b.device_ptrs()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    t1 = time.perf_counter()
    e2e_ms = allmax((t1 - t0) * 1e3) / args.e2e_steps
    h2d = int(sum(wl.input_bytes)) + 16 * nb
    d2h = int(sum(12 * V[k] + idx_bytes * I[k] + 16 * (int(wl.batches[k].sizes().num_meshes) + 1) for k in range(nb)))
    # the PCIe yardstick of the same run: a plain pinned D2H copy of the largest stream
    b = wl.batches[-1]
    b.sync()
    c0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        b._chk(b.L.vgb_batch_download(b.h, pos.ctypes.data, col.ctypes.data, idx.ctypes.data, len(pos), len(idx), None, None))
    c1 = time.perf_counter()
    copy_gbs = reps * (12 * V[-1] + idx_bytes * I[-1]) / (c1 - c0) / 1e9
    e2e = {"value": total_verts / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_ms, "d2h_gbs": d2h / (e2e_ms * 1e-3) / 1e9, "pinned_d2h_copy_gbs": copy_gbs,
           "pcie_frac": (d2h / (e2e_ms * 1e-3) / 1e9) / copy_gbs}
    dev_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        dev_step()
    t1 = time.perf_counter()
    dev_ms = allmax((t1 - t0) * 1e3) / args.e2e_steps
    e2e_dev = {"value": total_verts / (dev_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": dev_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 0,
               "what": "record + H2D + run + sync per step, mesh stays in HBM for a device-side consumer (vgb_batch_device_ptrs): a first run of a "
                       "fresh upload every step, no D2H of the mesh"}
    return e2e, e2e_dev


if __name__ == "__main__":
    sys.exit(main())
