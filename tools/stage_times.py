"""Per-stage device times of one resident 64-frame batch (bench workload). Usage: python tools/stage_times.py [frames]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, jxl_rs_b200 as j
from jxl_rs_b200 import abi
from concurrent.futures import ThreadPoolExecutor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = types.SimpleNamespace(frames=n, unique=0, width=3840, height=2160, distance=0.5, epf=2, profile=1)
files = bench.make_frames(args, 0)
with ThreadPoolExecutor(max_workers=16) as ex:
    frames = list(ex.map(j.ParsedFrame, files))
ctx = j.JxgContext(0)
outs = [torch.empty((fr.height, fr.width, 3), dtype=torch.uint8, device="cuda:0") for fr in frames]
b = j.Batch(ctx, n)
for fr, o in zip(frames, outs):
    b.add(fr, o.data_ptr(), fr.width * 3, abi.FORMAT_RGB_U8, True)
b.set_profile(True)
b.run(); b.wait()
best = None
for _ in range(4):
    b.rerun_device(); b.wait()
    st = b.stage_times()
    if best is None or sum(st.values()) < sum(best.values()):
        best = st
print({k: round(v, 2) for k, v in best.items() if v > 0.1}, "total", round(sum(best.values()), 2))
