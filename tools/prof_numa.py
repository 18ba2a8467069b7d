"""Which CPUs write pinned memory that the GPU's DMA engine then reads at full speed?"""
import os, sys, time, threading
import numpy as np, torch
dev = torch.device("cuda:0")
n = 21 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
hn = h.numpy(); src = np.random.randint(0, 255, n, dtype=np.uint8)
def dma():
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(3): dma()
print("main thread affinity", sorted(os.sched_getaffinity(0)))
res = {}
for c in sorted(os.sched_getaffinity(0)):
    def work():
        os.sched_setaffinity(0, {c})
        t0 = time.perf_counter(); np.copyto(hn, src); res[c] = (time.perf_counter() - t0) * 1e3
    th = threading.Thread(target=work); th.start(); th.join()
    print(f"cpu {c:2d}: copy {res[c]:.3f} ms, then DMA {dma():.3f} ms")
np.copyto(hn, src); print("main thread copy, then DMA %.3f ms" % dma())
