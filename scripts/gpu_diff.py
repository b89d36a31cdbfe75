"""Scratch: record-level diff of the device candidate records vs the oracle."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import checker as C
from dump1090_b200 import api, synth

def diff(name, data, aggressive):
    nbuf = data.size // api.BUFFER_BYTES + 1
    padded = np.full(nbuf * api.BUFFER_BYTES, 127, dtype=np.uint8); padded[:data.size] = data
    exp = C.oracle_scan_candidates(data, fix=1, aggressive=aggressive)
    exp_arr = np.frombuffer(b"".join(bytes(c) for c in exp), dtype=api.CANDIDATE_DTYPE)
    dec = api.Decoder(aggressive=aggressive)
    d = torch.from_numpy(padded).cuda()
    torch.cuda.synchronize()
    for rep in range(2):
        dec.detect_device(d.data_ptr(), nbuf)
        cands, tiles = dec.detect_fetch(nbuf)
        order = np.concatenate([np.arange(o, o + c) for o, c in tiles]).astype(int)
        got = cands[order]
        gb = got.view(np.uint8).reshape(-1, 56); eb = exp_arr.view(np.uint8).reshape(-1, 56)
        same_t = np.array_equal(got["t"], exp_arr["t"])
        bad = np.nonzero((gb != eb).any(axis=1))[0] if same_t else []
        print(name, "agg", aggressive, "rep", rep, "n", got.size, exp_arr.size, "t equal", same_t,
              "bad records", len(bad), "cols", sorted(set(np.nonzero(gb != eb)[1].tolist())) if same_t else "-")
        for i in list(bad)[:3]:
            print("  got", gb[i].tobytes().hex()); print("  exp", eb[i].tobytes().hex())
    dec.close()

diff("modes1", C.modes1(), 0)
diff("modes1", C.modes1(), 1)
diff("traffic1", synth.random_traffic(300000, 400, 1), 0)
diff("traffic1", synth.random_traffic(300000, 400, 1), 1)
