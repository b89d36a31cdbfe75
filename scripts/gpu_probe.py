"""Scratch GPU probe: record-level diff vs oracle + first timings. Not part of the product."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import checker as C
from dump1090_b200 import api, synth

def diff(data, aggressive):
    nbuf = data.size // api.BUFFER_BYTES + 1
    padded = np.full(nbuf * api.BUFFER_BYTES, 127, dtype=np.uint8); padded[:data.size] = data
    exp = C.oracle_scan_candidates(data, fix=1, aggressive=aggressive)
    exp_arr = np.frombuffer(b"".join(bytes(c) for c in exp), dtype=api.CANDIDATE_DTYPE)
    dec = api.Decoder(aggressive=aggressive)
    d = torch.from_numpy(padded).cuda()
    dec.detect_device(d.data_ptr(), nbuf)
    cands, tiles = dec.detect_fetch(nbuf)
    order = np.concatenate([np.arange(o, o + c) for o, c in tiles]).astype(int)
    got = cands[order]
    print("n", got.size, exp_arr.size, "t equal", np.array_equal(got["t"], exp_arr["t"]))
    gb = got.view(np.uint8).reshape(-1, 56); eb = exp_arr.view(np.uint8).reshape(-1, 56)
    bad = np.nonzero((gb != eb).any(axis=1))[0]
    print("mismatching records:", bad.size, "byte columns:", sorted(set(np.nonzero(gb != eb)[1].tolist())))
    for i in bad[:5]:
        print(i, "got", gb[i].tobytes().hex()); print(i, "exp", eb[i].tobytes().hex())
    dec.close()

diff(C.modes1(), 0)
diff(synth.random_traffic(300000, 400, 1), 1)

# first timing: 1 GiB tiled modes1, --no-fix
data = synth.tile_to(C.modes1(), 1 << 30)
nbuf = (1 << 30) // api.BUFFER_BYTES
d = torch.from_numpy(data).cuda()
dec = api.Decoder(fix_errors=0, profile=1)
for it in range(6):
    dec.detect_device(d.data_ptr(), nbuf)
    n = dec.detect_wait()
    t = dec.kernel_times_ms()
    print(f"iter {it}: cands {n}  scan {t[0]:.3f} ms ({2*(1<<29)/t[0]/1e6:.1f} GB/s)  eval {t[1]:.3f} ms  total {t[2]:.3f} ms -> {(1<<29)/t[2]/1e3:.1f} Msamples/s")
# e2e through process() from pinned memory
pb = api.PinnedBuffer(1 << 30); pb.array[:] = data
for it in range(3):
    dec.reset(); t0 = time.perf_counter(); dec.process_ptr(pb.ptr, 1 << 30); dec.finish(); t1 = time.perf_counter()
    print(f"e2e iter {it}: {t1-t0:.3f} s -> {(1<<29)/(t1-t0)/1e6:.1f} Msamples/s, msgs {len(dec.take_messages())}")
