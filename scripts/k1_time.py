"""Scratch: device time of the scan and frame-evaluation kernels over the 1 GiB bench workload."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench
from dump1090_b200 import api
cap, _ = bench.load_capture()
data = bench.shard_bytes(cap, 0, 1 << 30)
d = torch.from_numpy(np.ascontiguousarray(data)).cuda()
dec = api.Decoder(fix_errors=0, profile=1)
for _ in range(5):
    dec.detect_device(d.data_ptr(), 4096); dec.detect_wait()
dec.kernel_times_ms()
for _ in range(20):
    dec.detect_device(d.data_ptr(), 4096); n = dec.detect_wait()
t = dec.kernel_times_ms()
print(f"eval variant {os.environ.get('MODES_EVAL_VARIANT','fused')}: scan {t[0]:.4f} ms  eval {t[1]:.4f} ms  cands {n}  -> {2*(1<<29)/t[0]/1e6:.0f} GB/s", flush=True)
