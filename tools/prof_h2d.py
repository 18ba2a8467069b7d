import torch, time
dev = torch.device("cuda:0")
for mb in (1, 4, 21, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"pinned H2D {mb} MB: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"pinned D2H {mb} MB: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
import numpy as np
a = np.random.rand(3, 875000)
st = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
dst = st.numpy().view(np.float64).reshape(3, -1)
for _ in range(3): np.copyto(dst, a)
t0 = time.perf_counter()
for _ in range(10): np.copyto(dst, a)
print("host staging copy of 21 MB, 1 thread: %.3f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
