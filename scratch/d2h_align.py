"""D2H bandwidth into pinned memory vs destination alignment and copy size (proof download tuning)."""
import torch, time
dev = torch.device('cuda:0')
src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dst = torch.empty((1 << 30) + 8192, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
def run(nbytes, off, reps=8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        dst[off:off + nbytes].copy_(src[:nbytes], non_blocking=True)
        e0.record()
        for i in range(reps):
            o = off + i * nbytes
            dst[o:o + nbytes].copy_(src[:nbytes], non_blocking=True)
        e1.record()
    e1.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for nbytes in (64 << 20, 67108848, 8 << 20, 1 << 20):
    for off in (0, 8, 24, 72, 1000, 4096):
        print(f"size {nbytes:>10} dst offset {off:>5}: {run(nbytes, off, 8 if nbytes > (4 << 20) else 64):6.1f} GB/s")
