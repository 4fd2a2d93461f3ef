"""Host staging bandwidth probe: NumPy copy on 1..16 threads, torch copy_ (all CPU threads), pinned / pageable H2D — the numbers\nbehind predict()'s staging path (B200 host: 12.8 GB/s single-threaded, 50 GB/s with torch copy_, 37.7 GB/s pinned H2D)."""
import time, os, numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
a = np.random.rand(64, 512, 512, 3).astype(np.float32)
hp = torch.empty(a.shape, dtype=torch.float32).pin_memory()
h = hp.numpy()
pool = ThreadPoolExecutor(16)
def par(dst, src, nt):
    n = src.shape[0]; step = (n + nt - 1) // nt
    list(pool.map(lambda i: np.copyto(dst[i:i + step], src[i:i + step], casting="unsafe"), range(0, n, step)))
print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
for nt in (1, 2, 4, 8, 16):
    par(h, a, nt); t = time.time()
    for _ in range(5): par(h, a, nt)
    dt = (time.time() - t) / 5; print("numpy threads", nt, f"{dt*1e3:.1f} ms {a.nbytes/dt/1e9:.1f} GB/s")
ta = torch.from_numpy(a)
hp.copy_(ta); t = time.time()
for _ in range(5): hp.copy_(ta)
dt = (time.time() - t) / 5; print("torch copy_", f"{dt*1e3:.1f} ms {a.nbytes/dt/1e9:.1f} GB/s")
d = torch.empty(a.shape, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t = time.time()
for _ in range(5): d.copy_(hp, non_blocking=True)
torch.cuda.synchronize(); dt = (time.time() - t) / 5; print("H2D pinned", f"{dt*1e3:.1f} ms {a.nbytes/dt/1e9:.1f} GB/s")
torch.cuda.synchronize(); t = time.time()
for _ in range(3): d.copy_(ta)
torch.cuda.synchronize(); dt = (time.time() - t) / 3; print("H2D pageable", f"{dt*1e3:.1f} ms {a.nbytes/dt/1e9:.1f} GB/s")
