"""Seeded input streams shared by the CPU (host shim) and GPU parity tests: the demodulator's
nastiest inputs.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np

from dump1090_b200 import synth


def tie_rich(seed, n_frames=1500, levels=(0, 0, 1, 40, 100)):
    """Preambles followed by half-bit samples drawn from a few amplitudes: equal pairs (the
    reference's bits[0] == 2 and copied bits, dump1090.c:1675-1689), weak pairs, gate failures,
    all message types."""
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


def random_alphabet_cases(n_cases=24, seed=2024):
    """Small streams with random amplitude alphabets, preamble jitter, noise and saturation: weak
    and saturated pairs, long indefinite runs, every gate outcome."""
    rng = np.random.default_rng(seed)
    for case in range(n_cases):
        levels = tuple(int(x) for x in rng.choice([0, 0, 1, 2, 3, 5, 8, 20, 40, 70, 100, 127, 128], size=int(rng.integers(2, 6))))
        data = tie_rich(int(rng.integers(1, 1 << 30)), n_frames=int(rng.integers(60, 200)), levels=levels).copy()
        if case % 3 == 0:                                                 # additive noise on top
            noise = rng.integers(-3, 4, size=data.size)
            data = np.clip(data.astype(np.int64) + noise, 0, 255).astype(np.uint8)
        if case % 4 == 1:                                                 # saturated samples
            data[rng.integers(0, data.size, size=data.size // 50)] = 255
        yield case, levels, data


def retry_at_buffer_start(n_buffers=6, seed=3):
    """Frames whose preamble sits exactly at j == 0 of a reference buffer (stream sample
    131072k - 238) and whose first attempt is not a good message: the reference retries them
    WITHOUT phase correction (dump1090.c:1660).  Mix per buffer: unrepairable DF17 (3 flipped
    bits), DF17 with a flipped DF field, an address/parity frame of an unknown aircraft, a clean
    DF17 (control: no retry), each also one sample later (j == 1: corrected retry)."""
    rng = synth.Counter(seed)
    frames = []
    for k in range(1, n_buffers):
        j0 = 131072 * k - 238
        icao = 0x400000 + rng.below(0x3FFFFF)
        good = synth.make_frame(17, 5, icao.to_bytes(3, "big") + bytes(rng.below(256) for _ in range(7)))
        kind = k % 4
        if kind == 0:
            frame = synth.flip_bits(good, [20, 47, 81])
        elif kind == 1:
            frame = synth.flip_bits(good, [2, 60])
        elif kind == 2:
            frame = synth.make_frame(20, 3, bytes(rng.below(256) for _ in range(10)), icao_for_ap=0x123456 + k)
        else:
            frame = good
        frames.append((j0, frame, 70.0 + 5 * k, 0.4 * k, 0.3 if k % 2 else 0.0))
        frames.append((j0 + 65536 + 1, frame, 70.0, 0.2 * k, 0.4))          # far from any seam: normal corrected retry
    return synth.synth_stream(131072 * n_buffers + 3000, frames, sigma=1.2, seed=seed)
