"""Small Modular batch for ncu captures of k_modular_decode. Usage: python tools/ncu_modular.py [size=2048] [frames=4] [squeeze] [tree]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, synth, jxl_rs_b200 as j
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sq = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tk = int(sys.argv[4]) if len(sys.argv) > 4 else 1
files = [synth.encode_modular(size, size, 500 + i, 6, sq, tk) for i in range(n)]
frames = [j.ModularParsedFrame(f) for f in files]
ctx = j.JxgContext(0)
outs = [torch.empty((size, size, 3), dtype=torch.uint8, device="cuda:0") for _ in range(n)]
b = j.ModularBatch(ctx, 1)
for fr, o in zip(frames, outs):
    b.add(fr, o.data_ptr(), size * 3, True)
b.run(); b.wait()
print(b.stats())
b.close()
