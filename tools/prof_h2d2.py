import torch, time, numpy as np
dev = torch.device("cuda:0")
n = 21 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device=dev)
src = np.random.randint(0, 255, n, dtype=np.uint8)
hn = h.numpy()
def t_copy():
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for _ in range(3): t_copy()
print("untouched pinned buffer: %.3f ms" % t_copy())
for rep in range(3):
    np.copyto(hn, src)
    print("right after a CPU write (numpy copyto): %.3f ms" % t_copy(), "; again without write: %.3f ms" % t_copy())
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
except Exception as e: print(e)
