"""Scratch: e2e timing breakdown for different batch sizes."""
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bench
from dump1090_b200 import api
cap, _ = bench.load_capture()
pb = api.PinnedBuffer(1 << 30); pb.array[:] = bench.shard_bytes(cap, 0)
for mb in (16, 32, 64, 128, 256):
    dec = api.Decoder(fix_errors=0, max_batch_bytes=mb << 20)
    dec.set_output_array(700000)
    ts = []
    for it in range(5):
        dec.reset(); dec.rearm_output()
        t0 = time.perf_counter(); dec.process_ptr(pb.ptr, 1 << 30); dec.finish(); ts.append(time.perf_counter() - t0)
    print(f"batch {mb:4d} MiB: best {min(ts)*1e3:.2f} ms  median {sorted(ts)[2]*1e3:.2f} ms -> {(1<<29)/min(ts)/1e6:.0f} Msamples/s, msgs {dec.output_count()}", flush=True)
    dec.close()
os.environ["MODES_DEBUG_TIMING"] = "1"
dec = api.Decoder(fix_errors=0, max_batch_bytes=64 << 20); dec.set_output_array(700000)
t0 = time.perf_counter(); dec.process_ptr(pb.ptr, 1 << 30); dec.finish(); print("total", (time.perf_counter() - t0) * 1e3)
