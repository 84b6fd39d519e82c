"""Host-side probe on the GPU box: CPU quota and parse throughput of the front-end."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try:
        print(f, open(f).read().strip())
    except Exception as e:
        print(f, "n/a")
import bench, jxl_rs_b200 as j
args = types.SimpleNamespace(frames=64, unique=0, width=3840, height=2160, distance=0.5, epf=2, profile=1)
t = time.perf_counter(); files = bench.make_frames(args, 0); print("encode 64 frames", time.perf_counter() - t)
t = time.perf_counter(); fr = j.ParsedFrame(files[0]); print("one parse", time.perf_counter() - t)
for w in (8, 16, 32, 64, 128):
    with ThreadPoolExecutor(max_workers=w) as ex:
        t = time.perf_counter()
        frames = list(ex.map(j.ParsedFrame, files * 2))
        dt = time.perf_counter() - t
    print(f"workers={w}: 128 parses in {dt*1e3:.1f} ms -> {dt/128*1e3:.2f} ms/frame amortised")
    del frames
