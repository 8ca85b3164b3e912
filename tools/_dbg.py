import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
n = 1_000_000
b = vgb.Batch(0)
b.add_cubic_strokes(W.config3_cubics(n), width=4.0, cap=1, join=1)
b.upload()
mode = sys.argv[1] if len(sys.argv) > 1 else "sync"
em = []
for rep in range(3):
    for i in range(40):
        b.run()
        if mode == "sync":
            em.append(round(b.last_run_ms()["emit"], 3))
    b.sync()
    if mode != "sync":
        em.append(round(b.last_run_ms()["emit"], 3))
        for _ in range(4):
            b.run(); em.append(round(b.last_run_ms()["emit"], 3))
    time.sleep(0.2)
print(mode, em)
