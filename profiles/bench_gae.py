#!/usr/bin/env python
"""Throughput of the GAE kernel (mdx_gae, MAPPO/trainer.py:288-301) on a [T, N] rollout: CUDA events,
20 repetitions after 3 warm-ups.  Algorithmic bytes per element: rewards 4 + values 4 + dones 1 read,
advantages 4 + returns 4 written = 17 B.  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_delivery_b200 as mdb  # noqa: E402

T, N = 500, 65536
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
rew = torch.randn(T, N, device=dev, generator=g)
val = torch.randn(T, N, device=dev, generator=g)
done = (torch.rand(T, N, device=dev, generator=g) < 0.002).to(torch.uint8)
last = torch.randn(N, device=dev, generator=g)
for _ in range(3):
    mdb.compute_gae(rew, val, done, last)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    mdb.compute_gae(rew, val, done, last)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
peak = 6552.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gbs = 17.0 * T * N / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "mdx_gae_kernel", "T": T, "N": N, "ms_per_launch": ms, "elements_per_s": T * N / (ms * 1e-3),
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                               "bytes_per_element": 17}}))
