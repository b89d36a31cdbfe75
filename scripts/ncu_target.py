"""Short target for ncu captures: a few passes of the device hot path over the bench workload."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench
from dump1090_b200 import api
size = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
cap, _ = bench.load_capture()
data = bench.shard_bytes(cap, 0, size)
d = torch.from_numpy(np.ascontiguousarray(data)).cuda()
dec = api.Decoder(fix_errors=0)
for _ in range(3):
    dec.detect_device(d.data_ptr(), size // api.BUFFER_BYTES)
    print(dec.detect_wait())
