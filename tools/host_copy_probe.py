"""Host-side copy rates of a GPU box (decides how b200jk_upload / b200jk_fit_rows should feed the device):
threaded memcpy scaling, pageable H2D, cudaHostRegister cost, H2D out of registered memory."""
import json
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

out = {}
n = (1 << 30) // 8  # 1 GiB of doubles
src = np.ones(n)
dst = np.zeros(n)
for nt in (1, 2, 4, 8, 16):
    step = n // nt
    with ThreadPoolExecutor(nt) as ex:
        t0 = time.perf_counter()
        for _ in range(3):
            list(ex.map(lambda i: np.copyto(dst[i * step:(i + 1) * step], src[i * step:(i + 1) * step]), range(nt)))
        out[f"memcpy_{nt}_threads_GBs"] = 3 * n * 8 / (time.perf_counter() - t0) / 1e9
dev = torch.device("cuda", 0)
g = torch.empty(n, dtype=torch.float64, device=dev)
t = torch.from_numpy(src)
g.copy_(t)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    g.copy_(t)
torch.cuda.synchronize()
out["pageable_h2d_GBs"] = 3 * n * 8 / (time.perf_counter() - t0) / 1e9
rt = torch.cuda.cudart()
for gib in (1, 4):
    big = np.ones(gib * n)
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(big.ctypes.data, big.nbytes, 0)
    out[f"host_register_{gib}GiB_GBs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
    tb = torch.from_numpy(big)
    gg = torch.empty(gib * n, dtype=torch.float64, device=dev)
    gg.copy_(tb, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gg.copy_(tb, non_blocking=True)
    torch.cuda.synchronize()
    out[f"registered_h2d_{gib}GiB_GBs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    rt.cudaHostUnregister(big.ctypes.data)
    out[f"host_unregister_{gib}GiB_GBs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
    del gg, tb, big
t0 = time.perf_counter()
p = torch.empty(n // 2, dtype=torch.float64).pin_memory()
out["pinned_alloc_512MiB_s"] = time.perf_counter() - t0
print(json.dumps(out))
