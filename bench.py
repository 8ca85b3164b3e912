#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  metric   Mverts/s emitted (AA stroke), whole-job aggregate over all ranks
  workload configs[2]: 1M random cubic beziers per GPU (p0~U[0,1024)^2,
           p1..p3 = p0 + U[-128,128)^2), batched flatten + AA stroke, width 4, round caps + round
           joins, per-mesh uint16 indices.  Rank r takes items [r*1M, (r+1)*1M) of the stream
           (weak scaling, independent paths, no data-path collective).
  step     one pass of the hot path over the batch (vgb_batch_run: flatten -> count -> scan -> emit)
  value    inputs already resident in HBM; device time by CUDA events on the batch's own stream
           (event before step 0, event after step K-1: inter-step gaps and the run's internal
           size read-backs are included), max over ranks
  e2e      same metric through the C ABI with HOST buffers: record (memcpy of the pinned input
           into the batch tables) + H2D + run + D2H of pos/colour/index/mesh table into pinned
           host arrays, every step, wall clock between barriers, max over ranks
  roofline k_group_emit: algorithmic bytes (32*C + 12*V + 2*I) / its CUDA-event
           duration, against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline / --impl reference: the reference's own path.cpp + stroker.cpp (oracle/_ref,
           compiled from /root/reference against the bx shim) on the host cores

Each step writes ~3.9 GB of mesh per GPU, far more than the 126 MB L2, so no explicit L2 flush is
needed between steps (stated in config.l2).
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
# CPU reference arm (oracle/_ref): one process per host core, disjoint curve ranges
# ------------------------------------------------------------------------------------------
_BARRIER = None


def _cpu_init(barrier):
    global _BARRIER
    _BARRIER = barrier


def _cpu_worker(args):
    kind, first, count = args
    from oracle.pyref import RefLib
    from vg_renderer_b200 import workloads as W
    from vg_renderer_b200.tables import BatchTables
    lib = RefLib(kind)
    t = BatchTables.cubic_strokes(W.config3_cubics(count, first=first), width=WIDTH, cap=CAP, join=JOIN)
    lib.run_batch(t, store=False, path_range=(0, min(count, 2000)))  # warm caches / allocator
    _BARRIER.wait()  # every process starts the timed part together
    t0 = time.perf_counter()
    out = lib.run_batch(t, store=False)
    t1 = time.perf_counter()
    return out["num_vertices"], t0, t1


def cpu_reference_run(per_proc, procs, kind):
    """-> (total_vertices, wall_seconds): one process per core on disjoint ranges of the stream, all
    released by a barrier; wall = first start to last finish (CLOCK_MONOTONIC is system-wide)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(barrier,)) as pool:
        res = pool.map(_cpu_worker, [(kind, i * per_proc, per_proc) for i in range(procs)], chunksize=1)
    verts = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    return verts, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref_kind():
    from oracle.pyref import have_ref
    # stroke code is identical in the SIMD=0/1 builds; use the stock (SIMD=1) configuration
    return "sse" if have_ref("sse") else ("scalar" if have_ref("scalar") else None)


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    kind = ref_kind()
    cores = host_cores()
    per_proc = max(2000, min(100_000, args.ref_curves // max(cores, 1) if args.ref_curves else 50_000))
    if kind is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/*.so missing (make -C oracle ref needs /root/reference)"}))
        return 0
    times, verts = [], 0
    for s in range(args.warmup + args.steps):
        v, w = cpu_reference_run(per_proc, cores, kind)
        if s >= args.warmup:
            times.append(w)
            verts += v
    total = sum(times)
    value = verts / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1M random cubic beziers per GPU, batched flatten + AA stroke, width 4, round caps + round joins, uint16 indices",
                   "reference_sample": f"{cores} processes x {per_proc} cubics of the same stream per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{cores} procs x {per_proc} cubics/step, oracle/_ref/libvgref{'_sse' if kind == 'sse' else ''}.so (reference path.cpp+stroker.cpp, SIMD={'1' if kind == 'sse' else '0'})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """One process per GPU: run this rank on the CPUs next to its GPU, so that the pinned host buffers of
    the end-to-end leg (allocated later, first-touched by this process) live on the GPU's own NUMA node and
    the ranks' D2H streams do not all cross the socket interconnect.  Host-side placement only."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {"cpus": len(cpus), "first_cpu": cpus[0] if cpus else None}
    except Exception as e:  # no NVML / no permission: keep the default placement
        return {"error": str(e)[:80]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vgb", choices=["vgb", "reference"])
    ap.add_argument("--curves", type=int, default=CUBICS_PER_GPU, help="cubics per GPU (default: BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-curves", type=int, default=0, help="reference arm: total cubics per step (0 = 50k per core)")
    ap.add_argument("--merged", action="store_true", help="single merged uint32 index space: all-gather of per-rank (V,I) totals")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vgb" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist  # plumbing only: barrier + max-reduce of timings
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import vg_renderer_b200 as vgb
    from vg_renderer_b200 import workloads as W
    from vg_renderer_b200.api import pinned_empty

    if vgb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n = args.curves
    p8 = W.config3_cubics(n, first=rank * n)
    b = vgb.Batch(local)
    b.reset(1.0)
    b.add_cubic_strokes(p8, width=WIDTH, cap=CAP, join=JOIN)
    b.upload()
    vertex_base = 0
    if args.merged and dist is not None:
        # exclusive scan of per-rank vertex totals (16 B per rank over NCCL): SURVEY.md 8e
        from vg_renderer_b200.dist import exclusive_bases
        b.run()
        b.sync()
        s0 = b.sizes()
        vertex_base, _ib, _tv, _ti = exclusive_bases(int(s0.num_vertices), int(s0.num_indices), device=f"cuda:{local}")

    def step():
        b.run(index_u32=args.merged, vertex_base=vertex_base)

    for _ in range(args.warmup):
        step()
    b.sync()
    sampler = ClockSampler(local)
    barrier()
    b.sync()
    sampler.start()
    w0 = time.perf_counter()
    b.mark(0)
    emit_ms = []
    for _ in range(args.steps):
        step()
    b.mark(1)
    dev_ms = b.elapsed_ms(0, 1)
    b.sync()
    w1 = time.perf_counter()
    clocks = sampler.stop()
    barrier()
    sz = b.sizes()
    tm = b.last_run_ms()
    # per-kernel duration of the dominant kernel: a few extra (untimed) steps, CUDA events around emit
    for _ in range(5):
        step()
        emit_ms.append(b.last_run_ms()["emit"])
    emit_avg = float(np.mean(emit_ms))

    V, I, C = int(sz.num_vertices), int(sz.num_indices), n
    ms_per_step = allmax(dev_ms) / args.steps
    wall_ms_per_step = allmax((w1 - w0) * 1e3) / args.steps
    total_verts = allsum(float(V))
    value = total_verts / (ms_per_step * 1e-3) / 1e6

    idx_bytes = 4 if args.merged else 2
    alg_bytes = 32 * C + 12 * V + idx_bytes * I
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (emit_avg * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "emit_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("curves") == C and tj.get("index_bytes") == idx_bytes:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass
    roofline = {"bound": "hbm", "kernel": "k_group_emit", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": emit_avg,
                "pipeline_frac": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
                "stage_ms": {"flatten": tm["flatten"], "count_scan": tm["stroke_count"], "emit": tm["emit"], "total": tm["total"]}}

    # ---- e2e: host buffers in, host buffers out, every step --------------------------------
    e2e = None
    if not args.no_e2e:
        pos, _p1 = pinned_empty((V + 16, 2), np.float32)
        col, _p2 = pinned_empty((V + 16,), np.uint32)
        idx, _p3 = pinned_empty((I + 16,), np.uint32 if args.merged else np.uint16)
        pin_in, _p4 = pinned_empty((n, 8), np.float32)
        pin_in[:] = p8
        mvo, _p5 = pinned_empty((n + 1,), np.uint64)
        mio, _p6 = pinned_empty((n + 1,), np.uint64)

        def e2e_step():
            # host buffer in (borrowed, DMA'd by upload), host buffers out
            b.reset(1.0)
            b.set_cubic_strokes(pin_in, width=WIDTH, cap=CAP, join=JOIN)
            b.upload()
            b.run(index_u32=args.merged, vertex_base=vertex_base)
            b._chk(b.L.vgb_batch_download(b.h, pos.ctypes.data, col.ctypes.data, idx.ctypes.data, len(pos), len(idx),
                                          mvo.ctypes.data, mio.ctypes.data))

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        t1 = time.perf_counter()
        e2e_ms = allmax((t1 - t0) * 1e3) / args.e2e_steps
        h2d = 8 * 4 * n + 16  # the coordinates + one draw record (tables are generated on the device)
        d2h = 12 * V + idx_bytes * I + 16 * (n + 1)
        e2e = {"value": total_verts / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms, "d2h_gbs": d2h / (e2e_ms * 1e-3) / 1e9}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        kind = ref_kind()
        if kind is not None:
            cores = host_cores()
            per = 100_000
            v, wsec = cpu_reference_run(per, cores, kind)
            cpu = {"value": v / wsec / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{cores} procs x {per} cubics of the same stream, oracle/_ref/libvgref{'_sse' if kind == 'sse' else ''}.so, "
                             f"pathReset..strokerPolylineStrokeAA only, {wsec:.2f} s wall"}
            v1, w1s = cpu_reference_run(per, 1, kind)
            cpu["one_core_value"] = v1 / w1s / 1e6

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} random cubic beziers per GPU (SURVEY 8d config 3), batched flatten + AA stroke, width 4, round caps + round joins, "
                                   f"{'merged uint32' if args.merged else 'per-mesh uint16'} indices",
                       "verts_per_gpu": V, "indices_per_gpu": I, "points_per_gpu": int(sz.num_points),
                       "l2": "each step writes ~%.2f GB per GPU (>> 126 MB L2): no explicit flush" % (alg_bytes / 1e9),
                       "parallelism": f"{world} x independent path ranges, no data-path collective" + (" + 16 B/rank all-gather for the merged index base" if args.merged else ""),
                       "host_placement": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(sz.kernel_launches) * args.steps,
            "launches_per_step": int(sz.kernel_launches), "wall_ms_per_step": wall_ms_per_step,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
