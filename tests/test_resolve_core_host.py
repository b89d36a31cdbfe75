"""CPU check of the algorithm the device resolve runs (dump1090_b200/csrc/modes_resolve_core.cuh +
the buffer-parallel schedule of modes_resolve_gpu.cu): every reference buffer replayed on its own
from a guessed address cache, caches handed on slot by slot, wrong guesses repeated.  The host shim
(tests/host_shim/resolve_core_host.cpp, test infrastructure) must deliver exactly what the product's
sequential host resolver delivers, with the same statistics, on every stream and flag set."""
import ctypes
import subprocess

import numpy as np
import pytest

import checker as C
import streams as S
from dump1090_b200 import api, synth

SHIM = C.ROOT / "tests" / "_build" / "libresolve_core_host.so"


class Delivery(ctypes.Structure):
    _fields_ = [("t", ctypes.c_int64), ("pass_", ctypes.c_int32), ("crcok", ctypes.c_int32), ("phase_corrected", ctypes.c_int32),
                ("extra", ctypes.c_int32), ("extra_is_ap", ctypes.c_int32), ("pad", ctypes.c_int32)]


@pytest.fixture(scope="module")
def shim(checker_libs):
    subprocess.run(["make", "-s", "shim"], cwd=C.ROOT, check=True)
    lib = ctypes.CDLL(str(SHIM))
    lib.shim_resolve_buffers.restype = ctypes.c_long
    return lib


def _tiled(arr, n_buffers):
    g = (arr["t"] + 2) // api.TILE_SAMPLES
    nt = api.tiles_for(n_buffers)
    cnt = np.bincount(g, minlength=nt).astype(np.uint32)
    t = np.zeros(nt, dtype=api.TILE_DTYPE)
    t["count"] = cnt
    t["offset"] = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    return t


def _streams():
    yield "modes1", C.modes1()
    yield "traffic", synth.random_traffic(131072 * 9 + 5000, 1500, 77, n_aircraft=20)
    yield "dense_fleet", synth.random_traffic(131072 * 30, 9000, 5, n_aircraft=60)
    yield "grid", synth.df17_grid(280000, 700, 5)
    yield "ties", S.tie_rich(11)
    yield "retry_at_j0", S.retry_at_buffer_start()


STREAMS = dict(_streams())


@pytest.mark.parametrize("name", list(STREAMS))
@pytest.mark.parametrize("kw", [dict(), dict(aggressive=1), dict(check_crc=0), dict(check_crc=0, aggressive=1), dict(fix=0)],
                         ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()) or "default")
def test_buffer_parallel_resolve_equals_sequential(name, kw, shim):
    data = STREAMS[name]
    nbuf = data.size // api.BUFFER_BYTES + 1
    cands = C.oracle_scan_candidates(data, fix=kw.get("fix", 1), aggressive=kw.get("aggressive", 0), cap=400000)
    arr = np.frombuffer(b"".join(bytes(c) for c in cands), dtype=api.CANDIDATE_DTYPE).copy()
    tiles = _tiled(arr, nbuf)
    # the rule: the product's sequential host resolver
    seq = api.Resolver(fix_errors=kw.get("fix", 1), aggressive=kw.get("aggressive", 0), check_crc=kw.get("check_crc", 1))
    seq.run(arr, tiles)
    want = seq.take_messages()
    cap = 2 * len(cands) + 16
    out = (Delivery * cap)()
    stats = (ctypes.c_int64 * 8)()
    start = np.zeros(1024, dtype=np.uint32)
    end = np.zeros(1024, dtype=np.uint32)
    rounds = ctypes.c_int(0)
    n = shim.shim_resolve_buffers(arr.ctypes.data_as(ctypes.c_void_p), tiles.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.c_size_t(tiles.size), ctypes.c_size_t(nbuf), int(kw.get("check_crc", 1)),
                                  start.ctypes.data_as(ctypes.c_void_p), out, ctypes.c_size_t(cap), stats,
                                  end.ctypes.data_as(ctypes.c_void_p), ctypes.byref(rounds), api.TILE_SAMPLES)
    assert n == len(want)
    assert list(stats) == list(seq.stats().values())
    by_t = {int(c.t): c for c in cands}
    for d, m in zip(out[:n], want):
        ev = by_t[d.t].passes[d.pass_]
        assert d.t - 238 == m.sample_pos and bytes(ev.msg) == bytes(m.msg)
        assert (d.crcok, d.phase_corrected) == (m.crcok, m.phase_corrected)
        if d.extra_is_ap:
            assert d.extra == (m.aa1 << 16) | (m.aa2 << 8) | m.aa3
        else:
            assert d.extra == m.iid
    assert np.array_equal(end, seq.get_cache())
    assert 1 <= rounds.value <= nbuf
