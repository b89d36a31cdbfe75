"""CPU check of the packed arithmetic of the preamble scan kernel (dump1090_b200/csrc/modes_scan_core.cuh:
two squared magnitudes per register, comparisons as biased multiply-adds, flags gathered by byte
dot products).  The header is compiled for the host (tests/host_shim/, test infrastructure only)
and its masks are compared with a direct evaluation of dump1090.c:1602-1611 on the reference's
magnitudes.  The GPU parity tests check the kernel itself."""
import ctypes
import subprocess

import numpy as np
import pytest

import checker as C

SHIM = C.ROOT / "tests" / "_build" / "libscan_core_host.so"


@pytest.fixture(scope="module")
def shim():
    subprocess.run(["make", "-s", "shim"], cwd=C.ROOT, check=True)
    lib = ctypes.CDLL(str(SHIM))
    lib.shim_scan_masks.restype = ctypes.c_int
    lib.shim_scan_masks.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    lib.shim_scan_masks2.restype = ctypes.c_int
    lib.shim_scan_masks2.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    lib.shim_npack.restype = ctypes.c_uint32
    lib.shim_npack.argtypes = [ctypes.c_uint32]
    return lib


def _magnitudes(iq: np.ndarray) -> np.ndarray:
    """computeMagnitudeVector (dump1090.c:1454-1469) with the table formula of :359-364."""
    a = np.abs(iq.reshape(-1, 2).astype(np.int64) - 127)
    n = a[:, 0] ** 2 + a[:, 1] ** 2
    return np.floor(np.sqrt(n.astype(np.float64)) * 360 + 0.5).astype(np.int64)


def _expected(iq: np.ndarray, n_pos: int) -> np.ndarray:
    m = _magnitudes(iq)
    s = lambda d: m[d: d + n_pos]
    ok = ((s(0) > s(1)) & (s(1) < s(2)) & (s(2) > s(3)) & (s(3) < s(0)) & (s(4) < s(0)) & (s(5) < s(0))
          & (s(6) < s(0)) & (s(7) > s(8)) & (s(8) < s(9)) & (s(9) > s(6)))
    return ok


def _streams():
    rng = np.random.default_rng(5)
    n = 32 * 4096 + 64
    yield "noise", rng.normal(127, 3, 2 * n).round().clip(0, 255).astype(np.uint8)
    yield "uniform", rng.integers(0, 256, 2 * n, dtype=np.uint8)
    # few amplitude levels: ties everywhere, incl. the extremes 0 / 255 (|b-127| = 127 / 128)
    yield "levels", rng.choice(np.array([0, 1, 126, 127, 128, 254, 255], dtype=np.uint8), 2 * n)
    yield "saturated", rng.choice(np.array([255, 255, 255, 254, 127, 0], dtype=np.uint8), 2 * n)
    m1 = C.modes1()
    yield "modes1", m1[: (m1.size // 64) * 64]


@pytest.mark.parametrize("name,iq", list(_streams()), ids=[n for n, _ in _streams()])
def test_row_masks_match_direct_comparisons(shim, name, iq):
    n_samples = iq.size // 2
    n_pos = (n_samples - 16) // 32 * 32
    out = np.zeros(n_pos // 32, dtype=np.uint32)
    assert shim.shim_scan_masks(iq.ctypes.data, n_samples, n_pos, out.ctypes.data) == 0
    got = ((out[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).ravel()
    exp = _expected(iq, n_pos)
    assert exp.sum() > 0 or name == "saturated"
    assert np.array_equal(got, exp), f"{np.flatnonzero(got != exp)[:10]}"


@pytest.mark.parametrize("name,iq", list(_streams()), ids=[n for n, _ in _streams()])
def test_two_row_masks_match_direct_comparisons(shim, name, iq):
    """rows2_mask: the same comparisons with two rows (992 samples apart) sharing the registers."""
    n_samples = iq.size // 2
    row_b = 992
    n_pos = min((n_samples - row_b - 32) // 32 * 32, 32 * 600)
    a = np.zeros(n_pos // 32, dtype=np.uint32)
    b = np.zeros(n_pos // 32, dtype=np.uint32)
    assert shim.shim_scan_masks2(iq.ctypes.data, n_samples, n_pos, row_b, a.ctypes.data, b.ctypes.data) == 0
    bits = lambda m: ((m[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).ravel()
    exp = _expected(iq, row_b + n_pos)
    assert np.array_equal(bits(a), exp[:n_pos])
    assert np.array_equal(bits(b), exp[row_b: row_b + n_pos])


def test_npack_orders_like_the_magnitude_table(shim):
    """All 65536 (I, Q) byte pairs: the 15-bit field is order-equivalent to the reference magnitude."""
    i, q = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), indexing="ij")
    raw = (i | (q << 8)).ravel()
    lo = np.array([shim.shim_npack(int(r) | (0x7f7f << 16)) & 0xffff for r in raw])
    hi = np.array([shim.shim_npack((int(r) << 16) | 0x7f7f) >> 16 for r in raw])
    assert np.array_equal(lo, hi)
    mag = _magnitudes(np.stack([i.ravel(), q.ravel()], axis=1).astype(np.uint8).ravel())
    order = np.argsort(lo, kind="stable")
    assert lo.max() <= 0x7fff
    # equal fields <=> equal magnitudes, larger field <=> larger magnitude
    lo_s, mag_s = lo[order], mag[order]
    assert np.all((np.diff(lo_s) > 0) == (np.diff(mag_s) > 0))
    assert np.all(np.diff(mag_s) >= 0)
