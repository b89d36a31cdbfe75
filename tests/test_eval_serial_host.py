"""CPU check of the per-candidate evaluation that eval_serial_kernel runs one thread per candidate
(dump1090_b200/csrc/modes_eval_serial.cuh): the same header is compiled for the host
(tests/host_shim/, test infrastructure only) and fed the oracle's candidate positions; the
evaluated records — both attempts, frame bytes, gate/CRC/repair verdicts — must equal the
oracle's byte for byte.  The GPU parity tests check the kernel itself."""
import ctypes
import subprocess

import numpy as np
import pytest

import checker as C
from dump1090_b200 import api, synth

SHIM = C.ROOT / "tests" / "_build" / "libeval_serial_host.so"


@pytest.fixture(scope="module")
def shim(checker_libs):
    subprocess.run(["make", "-s", "shim"], cwd=C.ROOT, check=True)
    lib = ctypes.CDLL(str(SHIM))
    lib.shim_eval_candidates.restype = ctypes.c_int
    lib.shim_eval_candidates.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint32,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def _streams():
    yield "modes1", C.modes1()
    for seed in (1, 2, 3):
        yield f"traffic{seed}", synth.random_traffic(300000, 400, seed)
    yield "grid", synth.df17_grid(280000, 700, 5)
    yield "ties", _tie_rich(11)
    yield "ties_weak", _tie_rich(12, levels=(0, 0, 1, 3, 30))


def _tie_rich(seed, n_frames=1500, levels=(0, 0, 1, 40, 100)):
    """Preambles followed by half-bit samples drawn from a few amplitudes: equal pairs (the
    reference's bits[0] == 2 and copied bits), weak pairs, gate failures, all message types."""
    rng = np.random.default_rng(seed)
    period = 300
    iq = np.full((n_frames * period + 1000, 2), 127, dtype=np.uint8)
    for f in range(n_frames):
        s = 50 + f * period + int(rng.integers(0, 3))
        for k in (0, 2, 7, 9):
            iq[s + k, 0] = 127 + 110
        body = rng.choice(np.array(levels), size=224)
        if f % 4 == 0:
            body[1] = body[0]                                            # first pair ties
        iq[s + 16: s + 240, 0] = 127 + body
        iq[s + 16: s + 240, 1] = 127 - rng.choice(np.array(levels), size=224) // 2
    return iq.ravel()


STREAMS = dict(_streams())


def _evaluate(shim, data, fix, aggressive, lean):
    exp = C.oracle_scan_candidates(data, fix=fix, aggressive=aggressive, cap=400000)
    want = np.frombuffer(b"".join(bytes(c) for c in exp), dtype=api.CANDIDATE_DTYPE)
    nbuf = data.size // api.BUFFER_BYTES + 1
    virt = np.full(480 + nbuf * api.BUFFER_BYTES, 127, dtype=np.uint8)   # halo (initial carry: no signal) + whole buffers
    virt[480: 480 + data.size] = data
    v = (want["t"] + 2).astype(np.uint32)
    got = np.zeros(want.size, dtype=api.CANDIDATE_DTYPE)
    rc = shim.shim_eval_candidates(virt.ctypes.data, virt.size // 2, v.ctypes.data, v.size, fix, aggressive, lean,
                                   got.ctypes.data)
    assert rc == 0
    return got.view(np.uint8).reshape(-1, 56), want.view(np.uint8).reshape(-1, 56)


@pytest.mark.parametrize("name", list(STREAMS))
@pytest.mark.parametrize("fix,aggressive", [(1, 0), (1, 1), (0, 0)])
@pytest.mark.parametrize("lean", [0, 1], ids=["default", "lean"])
def test_serial_evaluation_matches_oracle(name, fix, aggressive, lean, shim):
    got, want = _evaluate(shim, STREAMS[name], fix, aggressive, lean)
    assert want.shape[0] > 200
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} of {want.shape[0]} records differ, first at {bad[0]}: {got[bad[0]].tolist()} != {want[bad[0]].tolist()}"


@pytest.mark.parametrize("lean", [0, 1], ids=["default", "lean"])
def test_serial_evaluation_random_alphabets(lean, shim):
    """Many small streams with random amplitude alphabets, preamble jitter and noise levels: weak
    and saturated pairs, long indefinite runs, every gate outcome."""
    rng = np.random.default_rng(2024)
    total = 0
    for case in range(24):
        levels = tuple(int(x) for x in rng.choice([0, 0, 1, 2, 3, 5, 8, 20, 40, 70, 100, 127, 128], size=int(rng.integers(2, 6))))
        data = _tie_rich(int(rng.integers(1, 1 << 30)), n_frames=int(rng.integers(60, 200)), levels=levels).copy()
        if case % 3 == 0:                                                 # additive noise on top
            noise = rng.integers(-3, 4, size=data.size)
            data = np.clip(data.astype(np.int64) + noise, 0, 255).astype(np.uint8)
        if case % 4 == 1:                                                 # saturated samples
            data[rng.integers(0, data.size, size=data.size // 50)] = 255
        for fix, aggressive in ((1, 1), (0, 0)):
            got, want = _evaluate(shim, data, fix, aggressive, lean)
            total += want.shape[0]
            bad = np.nonzero((got != want).any(axis=1))[0]
            assert bad.size == 0, f"case {case} levels {levels}: {bad.size} of {want.shape[0]} records differ"
    assert total > 5000
