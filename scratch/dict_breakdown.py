import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_delivery_b200 as mdb
from marl_delivery_b200.batch import encode_action
def us(fn, n=2000):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
e = mdb.Environment("map1.txt", 500, 5, 50, seed=42); e.reset()
b = e._b
acts = [("D", "1"), ("L", "0"), ("S", "2"), ("U", "0"), ("R", "1")]
codes = np.array([[encode_action(m, p) for (m, p) in acts]], dtype=np.uint8)
out = {}
out["encode+array"] = us(lambda: np.array([[encode_action(m, p) for (m, p) in acts]], dtype=np.uint8))
out["step_host (C call incl. numpy allocs)"] = us(lambda: b.step_host(codes, [0]))
r = b.step_host(codes, [0])[3]
out["rows_to_states"] = us(lambda: b._rows_to_states(r))
e._seed(b._rows_to_states(r)[0])
out["get_state dict"] = us(lambda: e.get_state())
out["Environment.step total"] = us(lambda: e.step(acts))
import ctypes as C
reward = np.empty(1, np.float64); done = np.empty(1, np.uint8); rows = np.empty(1, dtype=b._row_dtype); h = np.array([0], np.int32)
L = b._L
out["raw mdx_step_host ctypes"] = us(lambda: L.mdx_step_host(b._h, codes.ctypes.data, h.ctypes.data, 1, reward.ctypes.data, None, done.ctypes.data, rows.ctypes.data, 0, None))
out["raw mdx_step_host no rows"] = us(lambda: L.mdx_step_host(b._h, codes.ctypes.data, h.ctypes.data, 1, reward.ctypes.data, None, done.ctypes.data, None, 0, None))
out["raw mdx_get_states"] = us(lambda: L.mdx_get_states(b._h, h.ctypes.data, 1, rows.ctypes.data, None))
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
