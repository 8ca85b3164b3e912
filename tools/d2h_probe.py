"""D2H bandwidth of pinned buffers vs CPU affinity / number of streams (tuning helper)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

def numa_of_gpu(i=0):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
    bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
    if len(bdf.split(":")[0]) == 8: bdf = bdf[4:]
    try:
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except Exception as e:
        return None, str(e)
    cpus = open(f"/sys/devices/system/node/node{max(node,0)}/cpulist").read().strip()
    return node, cpus

print("gpu0 numa:", numa_of_gpu(0), "affinity now:", len(os.sched_getaffinity(0)), "cpus")
print(open("/proc/cpuinfo").read().count("processor"), "logical cpus;", os.popen("lscpu | grep -E 'NUMA|Socket|Model name'").read())
dev = torch.device("cuda:0")
n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, device=dev)
def bw(label, streams=1):
    dst = torch.empty(n, dtype=torch.uint8).pin_memory()
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize()
    best = 0
    for _ in range(3):
        t0 = time.perf_counter()
        chunk = n // streams
        for k, s in enumerate(ss):
            with torch.cuda.stream(s):
                dst[k*chunk:(k+1)*chunk].copy_(src[k*chunk:(k+1)*chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    print(f"{label:40s} streams={streams}: {best:.1f} GB/s")
bw("default affinity", 1); bw("default affinity", 3)
node, cpus = numa_of_gpu(0)
if node is not None and node >= 0:
    def parse(c):
        out = []
        for p in c.split(","):
            a, _, b = p.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    for nd in sorted({int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node")}):
        cl = parse(open(f"/sys/devices/system/node/node{nd}/cpulist").read().strip())
        try:
            os.sched_setaffinity(0, cl)
            bw(f"affinity node{nd}{' (gpu-local)' if nd == node else ''}", 1)
        except Exception as e:
            print("node", nd, e)
