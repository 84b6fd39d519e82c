"""Sweeps the persistent-lane knobs (JXG_ENTROPY_S / JXG_ENTROPY_LANES) of k_entropy_lean on one resident batch.
Usage: python tools/sweep_entropy.py [frames]"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, jxl_rs_b200 as j
from jxl_rs_b200 import abi
from concurrent.futures import ThreadPoolExecutor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
args = types.SimpleNamespace(frames=n, unique=0, width=3840, height=2160, distance=0.5, epf=2, profile=1)
files = bench.make_frames(args, 0)
with ThreadPoolExecutor(max_workers=64) as ex:
    frames = list(ex.map(j.ParsedFrame, files))
ctx = j.JxgContext(0)
outs = [torch.empty((fr.height, fr.width, 3), dtype=torch.uint8, device="cuda:0") for fr in frames]
def build():
    b = j.Batch(ctx, n)
    for fr, o in zip(frames, outs):
        b.add(fr, o.data_ptr(), fr.width * 3, abi.FORMAT_RGB_U8, True)
    b.set_profile(True)
    b.run(); b.wait()
    return b
def measure(tag):
    b = build()
    best = None
    for _ in range(3):
        b.rerun_device(); b.wait()
        st = b.stage_times()
        if best is None or st["entropy"] < best["entropy"]:
            best = st
    b.close()
    print(tag, {k: round(v, 2) for k, v in best.items() if v > 0.1}, flush=True)
measure("default")
for S in (2, 4, 8):
    for M in (1.0, 1.3, 1.6, 2.0):
        os.environ["JXG_ENTROPY_S"] = str(S)
        os.environ["JXG_ENTROPY_LANES_MUL"] = str(M)
        measure(f"S={S} mul={M}")
