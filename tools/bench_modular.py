"""BASELINE config 5 measurement: a batch of synthetic lossless 4096x4096 Modular frames on one GPU.
Usage: python tools/bench_modular.py [frames=8] [squeeze=0] [tree_kind=1] [lanes=1,2,4]
Prints one JSON line per lanes setting (device-resident MP/s, decode-kernel ms, CPU checker MP/s on a 1-frame sample)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch
import synth, jxl_rs_b200 as j

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
squeeze = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tree_kind = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lanes_list = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,2,4").split(",")]
W = H = 4096
t = time.perf_counter()
with ThreadPoolExecutor(max_workers=min(n, 16)) as ex:
    files = list(ex.map(lambda s: synth.encode_modular(W, H, 500 + s, 6, squeeze, tree_kind), range(n)))
t_enc = time.perf_counter() - t
t = time.perf_counter()
with ThreadPoolExecutor(max_workers=min(n, 16)) as ex:
    frames = list(ex.map(j.ModularParsedFrame, files))
t_parse = time.perf_counter() - t
ctx = j.JxgContext(0)
outs = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda:0") for _ in range(n)]
for lanes in lanes_list:
    b = j.ModularBatch(ctx, lanes)
    for fr, o in zip(frames, outs):
        b.add(fr, o.data_ptr(), W * 3, True)
    b.run(); b.wait()
    best = None
    for _ in range(3):
        b.rerun_device(); b.wait()
        st = b.stats()
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
    ok = bool(np.array_equal(outs[0].cpu().numpy(), synth.modular_source(W, H, 500)))
    b.close()
    print(json.dumps({"workload": f"{n} x {W}x{H} lossless Modular, RCT 6, squeeze {squeeze}, tree kind {tree_kind}", "lanes_per_warp": lanes,
                      "device_ms": round(best["device_ms"], 2), "decode_kernel_ms": round(best["decode_ms"], 2),
                      "mp_per_s_device": round(n * W * H / 1e6 / (best["device_ms"] / 1e3)), "bit_exact_vs_source": ok,
                      "bytes_per_frame": len(files[0]), "host_parse_s": round(t_parse, 2), "encode_s": round(t_enc, 1),
                      "kernel_launches": best["kernel_launches"]}), flush=True)
from tests import oracle_binding as ob
t = time.perf_counter(); ob.decode_modular_file(files[0]); t_cpu = time.perf_counter() - t
print(json.dumps({"cpu_checker_1_thread_mp_per_s": round(W * H / 1e6 / t_cpu, 1), "sample": "1 frame"}))
