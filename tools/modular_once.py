"""One resident Modular batch of the config 5 shape, decoded a few times (for ncu captures and quick timings).
Usage: python tools/modular_once.py [frames] [size] [tree_kind] [reruns]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jxl_rs_b200 as j
import synth
from concurrent.futures import ThreadPoolExecutor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
tk = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reruns = int(sys.argv[4]) if len(sys.argv) > 4 else 3
with ThreadPoolExecutor(max_workers=8) as ex:
    files = list(ex.map(lambda sd: synth.encode_modular(size, size, sd, 6, 0, tk), range(500, 500 + n)))
    frames = list(ex.map(j.ModularParsedFrame, files))
ctx = j.JxgContext(0)
outs = [torch.empty((size, size, 3), dtype=torch.uint8, device="cuda:0") for _ in range(n)]
b = j.ModularBatch(ctx, int(os.environ.get("MODULAR_LANES", "1")))
for fr, o in zip(frames, outs):
    b.add(fr, o.data_ptr(), size * 3, True)
b.run(); b.wait()
best = None
for _ in range(reruns):
    b.rerun_device(); b.wait()
    st = b.stats()
    if best is None or st["device_ms"] < best["device_ms"]:
        best = st
print(f"{n} x {size}x{size} tree_kind {tk}: device {best['device_ms']:.2f} ms, decode kernel {best['decode_ms']:.2f} ms, "
      f"{n*size*size/1e6/(best['device_ms']/1e3):.0f} MP/s")
b.close(); ctx.close()
