"""Small decode under compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import checker as C
from dump1090_b200 import api, synth
data = synth.random_traffic(131072 * 2 + 3000, 300, 3)
pat = np.array([1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0], dtype=np.uint8)
dense = np.full(2 * (131072 + 999), 127, dtype=np.uint8)
dense[0::2] = 127 + 100 * np.tile(pat, -(-(131072 + 999) // 15))[: 131072 + 999]
for name, d, kw, okw in [("traffic", data, dict(aggressive=1), dict(aggressive=1)),
                         ("dense", dense, dict(check_crc=0), dict(check_crc=0, cap=200000)),
                         ("traffic, device resolve", data, dict(aggressive=1, gpu_resolve=1, max_batch_bytes=262144), dict(aggressive=1)),
                         ("traffic, 2 GPU shards", data, dict(n_gpus=2, max_batch_bytes=262144), dict())]:
    exp, st = C.oracle_decode(d, **okw)
    with api.Decoder(**kw) as dec:
        got = dec.decode(d)
        ok = [m.raw_line() for m in got] == [m.hexline() for m in exp] and list(dec.stats().values()) == st
    print(name, len(got), "parity", "OK" if ok else "MISMATCH", flush=True)
