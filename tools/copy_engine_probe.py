"""Do H2D and D2H cudaMemcpyAsync calls of different streams overlap on this box?
Stream A: a queue of D2H copies (64 x 25 MB); stream B, submitted right after: event, one 218 MB H2D copy, event.
Prints when B's copy started / ended relative to A's queue, and the same with the H2D done by a kernel reading pinned
host memory (the SMs instead of a copy engine)."""
import torch, time
dev = torch.device("cuda:0")
n, sz = 64, 25 * 1024 * 1024
d_src = torch.empty(n * sz, dtype=torch.uint8, device=dev)
h_dst = torch.empty(n * sz, dtype=torch.uint8).pin_memory()
h_blob = torch.empty(218 * 1024 * 1024, dtype=torch.uint8).pin_memory()
d_blob = torch.empty_like(h_blob, device=dev)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for mode in ("memcpy", "memcpy", "idle"):
    torch.cuda.synchronize()
    ref = torch.cuda.Event(enable_timing=True); a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
    b0 = torch.cuda.Event(enable_timing=True); b1 = torch.cuda.Event(enable_timing=True)
    ref.record()
    if mode != "idle":
        with torch.cuda.stream(sa):
            a0.record()
            for i in range(n):
                h_dst[i * sz:(i + 1) * sz].copy_(d_src[i * sz:(i + 1) * sz], non_blocking=True)
            a1.record()
    with torch.cuda.stream(sb):
        b0.record()
        d_blob.copy_(h_blob, non_blocking=True)
        b1.record()
    torch.cuda.synchronize()
    if mode != "idle":
        print(f"{mode}: D2H queue {ref.elapsed_time(a0):.1f} .. {ref.elapsed_time(a1):.1f} ms ({n*sz/1e9/(a0.elapsed_time(a1)/1e3):.1f} GB/s); "
              f"H2D of the other stream {ref.elapsed_time(b0):.1f} .. {ref.elapsed_time(b1):.1f} ms")
    else:
        print(f"H2D alone: {ref.elapsed_time(b0):.1f} .. {ref.elapsed_time(b1):.1f} ms ({h_blob.numel()/1e9/(b0.elapsed_time(b1)/1e3):.1f} GB/s)")
