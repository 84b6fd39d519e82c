"""Host-side timeline of PipelinedDecoder on the bench workload.
Usage: python tools/e2e_profile4.py [frames] [steps] [depth]; E2E_SPLIT=k submits every step as k sub-batches."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, jxl_rs_b200 as j
from jxl_rs_b200 import abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
args = types.SimpleNamespace(frames=n, unique=0, width=3840, height=2160, distance=0.5, epf=2, profile=1)
files = bench.make_frames(args, 0)
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
j.bind_to_gpu_numa_node(0)
split = int(os.environ.get("E2E_SPLIT", "1"))
sets = (depth + split - 1) // split + 1  # host output sets: a set is rewritten only after its sub-batches retired
host_out = [[torch.empty((2160, 3840, 3), dtype=torch.uint8).pin_memory() for _ in range(n)] for _ in range(sets)]
outs = [[(o.data_ptr(), 3840 * 3) for o in ho] for ho in host_out]
per = n // split


def submit(dec, i):
    for s in range(split):
        dec.submit(files[s * per:(s + 1) * per], outs[i % sets][s * per:(s + 1) * per])


for st in [int(x) for x in os.environ.get('E2E_STAGING', '4,8').split(',')]:
    dec = j.PipelinedDecoder(0, depth=depth, staging_threads=st)
    if os.environ.get("E2E_MARKS"):
        dec.marks = []  # the warm-up batches set the zero of the device timeline (before the timed region)
    for i in range(depth + 1):  # every context has sized its pools before the clock starts
        submit(dec, i)
    dec.drain()
    # reference copy of the last rows of the last frame (from the warm-up), then poison those rows in every output set:
    # after drain() they must hold the picture again - i.e. drain() really returns with the pixels in host memory
    ref_rows = host_out[0][n - 1][-64:].clone()
    for ho in host_out:
        ho[n - 1][-64:].zero_()
    dec.trace = []
    dec.retire_trace = []
    dec.marks = [] if os.environ.get("E2E_MARKS") else None
    t0 = time.perf_counter()
    for i in range(steps):
        submit(dec, i)
    t_sub = time.perf_counter()
    dec.drain()
    t_drain = time.perf_counter()
    complete = bool(torch.equal(host_out[(steps - 1) % sets][n - 1][-64:], ref_rows))  # before any device-wide wait
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_sync2 = time.perf_counter()
    print(f"  second synchronize: {1e3*(t_sync2-t0-dt):.2f} ms")
    print(f"staging_threads={st}: {dt/steps*1e3:.1f} ms/step, {n*3840*2160/1e6*steps/dt:.0f} MP/s")
    print(f"  main thread: submits done at {1e3*(t_sub-t0):.1f} ms, drain returned at {1e3*(t_drain-t0):.1f} "
          f"(last rows of the last frame in host memory: {complete}), device idle at {1e3*dt:.1f}")
    for (ts, ret, pw, add, run) in dec.trace:
        print(f"  t={1e3*(ts-t0):7.1f}  retire {ret*1e3:6.1f}  parsewait {pw*1e3:6.1f}  add {add*1e3:6.1f}  run {run*1e3:6.1f}")
    for (ts, w, m, c) in dec.retire_trace:
        print(f"  retire at t={1e3*(ts-t0):7.1f}: wait {w*1e3:6.1f}  stats+marks {m*1e3:6.1f}  close {c*1e3:6.1f}")
    if dec.marks:
        # device times on the host clock of the lines above: zero of the device timeline = dec.marks_ref_host
        base = -1e3 * (dec.marks_ref_host - t0)
        print("  device timeline (ms): start | entropy begin..end | transforms end | filters launches end | H2D start | D2H end")
        for m in dec.marks:
            print("   ", " ".join(f"{(v - base):8.1f}" if v != 0 else "       -" for v in m))
    dec.close()
