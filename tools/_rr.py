import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vg_renderer_b200 as vgb
from vg_renderer_b200 import workloads as W
r = int(sys.argv[1]) if len(sys.argv) > 1 else 512
b = vgb.Batch(0)
b.add_rounded_rect_fills(W.config4_rrects(1_000_000, r))
for _ in range(4):
    b.run()
print(r, b.last_run_ms())
