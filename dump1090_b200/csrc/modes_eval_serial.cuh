// modes_eval_serial.cuh — evaluation of ONE preamble candidate by ONE thread.
//
// This is the body of detectModeS after the preamble test (dump1090.c:1653-1735: bit slicing, delta
// gate, the phase-corrected retry of :1498-1558) and the order-independent half of
// decodeModesMessage (:1099-1128: CRC syndrome, single/two-bit repair), written as straight
// sequential code over the candidate's 241-sample window.  eval_serial_kernel (modes_kernels.cu)
// runs it with lane = candidate, the 32 windows of a warp staged in shared memory.
//
// The file has no CUDA dependencies besides a few integer intrinsics, so the test suite also
// compiles it for the host (tests/host_shim/) and checks the logic against the oracle without a
// GPU.  That build is test infrastructure; the product library contains only the device code.
#pragma once
#include <cstdint>
#include "modes_b200.h"

#if defined(__CUDACC__)
#define MODES_SERIAL_FN __device__ __forceinline__
#else
#define MODES_SERIAL_FN static inline
#endif

namespace modes {
namespace serial {

// ---- the intrinsics ----------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MODES_SERIAL_FN uint32_t absdiff127x4(uint32_t w) { return __vabsdiffu4(w, 0x7f7f7f7fu); }
MODES_SERIAL_FN uint32_t dot4(uint32_t a, uint32_t b) { return __dp4a(a, b, 0u); }
MODES_SERIAL_FN uint32_t bitrev(uint32_t x) { return __brev(x); }
// (hi:lo) >> s for s in {16, 32}
MODES_SERIAL_FN uint32_t funnel(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_rc(lo, hi, s); }
MODES_SERIAL_FN uint32_t bswap(uint32_t x) { return __byte_perm(x, 0u, 0x0123u); }
MODES_SERIAL_FN uint32_t absdiff_add(uint32_t a, uint32_t b, uint32_t c) { return __usad(a, b, c); }   // |a-b| + c
MODES_SERIAL_FN uint32_t perm(uint32_t x, uint32_t sel) { return __byte_perm(x, 0u, sel); }             // bytes of x picked by sel's nibbles
MODES_SERIAL_FN uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }                    // keep a value in its register
#else
MODES_SERIAL_FN uint32_t bswap(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }
MODES_SERIAL_FN uint32_t absdiff_add(uint32_t a, uint32_t b, uint32_t c) { return (a > b ? a - b : b - a) + c; }
MODES_SERIAL_FN uint32_t perm(uint32_t x, uint32_t sel) {
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        const uint32_t n = (sel >> (4 * k)) & 7u;
        r |= (n < 4 ? (x >> (8 * n)) & 0xffu : 0u) << (8 * k);
    }
    return r;
}
MODES_SERIAL_FN uint32_t opaque(uint32_t x) { return x; }
MODES_SERIAL_FN uint32_t absdiff127x4(uint32_t w) {
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        int b = (int)((w >> (8 * k)) & 0xff) - 127;
        r |= (uint32_t)(b < 0 ? -b : b) << (8 * k);
    }
    return r;
}
MODES_SERIAL_FN uint32_t dot4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r += ((a >> (8 * k)) & 0xff) * ((b >> (8 * k)) & 0xff);
    return r;
}
MODES_SERIAL_FN uint32_t bitrev(uint32_t x) {
    uint32_t r = 0;
    for (int k = 0; k < 32; k++) r |= ((x >> k) & 1u) << (31 - k);
    return r;
}
MODES_SERIAL_FN uint32_t funnel(uint32_t lo, uint32_t hi, uint32_t s) {
    return s >= 32 ? hi : (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
}
#endif

// ---- window layout ----------------------------------------------------------------------------
// A candidate at virtual position v owns the samples m[-1..239] = virtual samples v-1 .. v+239.
// They are staged as kWindowWords aligned 32-bit words (two samples each); `odd` says whether
// m[-1] is the high half of word 0.  Window sample w (w = 0 is m[-1]) is halfword w + odd.
// Bit b of the frame is the sample pair (m[16+2b], m[17+2b]) = window samples 17+2b, 18+2b, i.e.
// halfwords of words 8+b and 9+b.  After the first pass word 8+b holds the pair's magnitudes
// (first | second << 16) instead; words 0..7 keep the raw preamble samples.
constexpr int kWindowWords = 121;

// Rows of the (|I-127|, |Q-127|) magnitude table are 136 entries (68 words) apart: consecutive rows
// start 4 banks apart in shared memory, so the small amplitudes that most half-bit samples have
// (|I-127|, |Q-127| < 8) spread over all 32 banks instead of piling onto two or three.
constexpr int kIqLutStride = 136;
constexpr int kIqLutEntries = 129 * kIqLutStride;

struct Tables {
    const uint16_t *lut_iq;      // [129 rows of kIqLutStride] magnitude by (|I-127|, |Q-127|): round(sqrt(i*i+q*q)*360), dump1090.c:362
    const uint32_t *bit_syn;     // [112]   syndrome of one flipped bit
    const uint32_t *nib_syn;     // [28*16] syndrome of nibble value x at frame nibble i: XOR of bit_syn over its set bits
    const uint32_t *fix_hash;    // [256]   inverse of bit_syn over positions 5..111
    const uint32_t *pair_hash;   // [1 << kPairHashBits] syndrome of two flipped bits p < q -> p (global memory on the device)
};
constexpr int kPairHashBits = 14, kPairHashMaxProbe = 16;

// First flipped bit p of the two-bit pattern with syndrome s, or -1.
MODES_SERIAL_FN int pair_first(const uint32_t *pair_hash, uint32_t s) {
    const uint32_t h = (s * 0x9E3779B1u) >> (32 - kPairHashBits);
    for (int i = 0; i < kPairHashMaxProbe; i++) {
        const uint32_t e = pair_hash[(h + i) & ((1u << kPairHashBits) - 1u)];
        if (e == 0xFFFFFFFFu) return -1;
        if ((e >> 7) == s) return (int)(e & 0x7f);
    }
    return -1;
}

// Index into lut_iq of the sample held in the low (kLutLow) or high (kLutHigh) half of a word of
// |byte - 127| values: kIqLutStride * |I-127| + |Q-127| as one byte dot product.
constexpr uint32_t kLutLow = (uint32_t)kIqLutStride | 0x100u, kLutHigh = kLutLow << 16;

// What one attempt (uncorrected, or phase corrected) yields: the six 32-bit words of a
// modes_frame_eval (include/modes_b200.h) are made from it by eval_words().
struct Verdict {
    uint32_t F[4];               // frame bits, bit b of the frame at bit b (LSB first); after repair
    uint32_t msgtype, flags, errorbit, nfixed, crc;
};

// Entry (pos, value) of Tables::nib_syn from the bit syndromes: bit 3-k of the nibble is frame bit 4*pos+k.
MODES_SERIAL_FN uint32_t nibble_syndrome(const uint32_t *bit_syn, int pos, uint32_t value) {
    uint32_t x = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) x ^= ((value >> (3 - k)) & 1u) ? bit_syn[4 * pos + k] : 0u;
    return x;
}

MODES_SERIAL_FN uint32_t fix_hash_of(uint32_t s) { return (s * 0x9E3779B1u) >> 24; }

// Position (5..111) whose single-bit syndrome is s, or -1.
MODES_SERIAL_FN int syndrome_pos(const uint32_t *fix_hash, uint32_t s) {
    uint32_t h = fix_hash_of(s);
    for (int i = 0; i < 256; i++) {
        uint32_t e = fix_hash[(h + i) & 255u];
        if (e == 0xFFFFFFFFu) return -1;
        if ((e >> 8) == s) return (int)(e & 0xff);
    }
    return -1;
}

MODES_SERIAL_FN void flip_bit(uint32_t F[4], int b) {
    // no dynamic register indexing: four predicated XORs
    const uint32_t m = 1u << (b & 31);
    const int w = b >> 5;
    F[0] ^= (w == 0) ? m : 0u; F[1] ^= (w == 1) ? m : 0u;
    F[2] ^= (w == 2) ? m : 0u; F[3] ^= (w == 3) ? m : 0u;
}

// Frame nibble i (0..27) of F: bit 4i is its most significant bit.
MODES_SERIAL_FN uint32_t frame_nibble(const uint32_t W[4], int i) {      // W = bit-reversed F words
    const uint32_t w = (i < 8) ? W[0] : (i < 16) ? W[1] : (i < 24) ? W[2] : W[3];
    return (w >> (28 - 4 * (i & 7))) & 0xfu;
}

// CRC syndrome of the first msgbits of F, then the repairs of dump1090.c:1114-1126 (:733-742
// checksum, :854-894 fixes).  F is modified by a repair.
MODES_SERIAL_FN void crc_and_fix(uint32_t F[4], int msgbits, uint32_t msgtype, int fix_errors, int aggressive,
                                 const Tables &tab, uint32_t &crc, uint32_t &errorbit, uint32_t &nfixed) {
    const int off = 112 - msgbits;                      // a short frame uses the last 56 table positions
    const uint32_t W[4] = {bitrev(F[0]), bitrev(F[1]), bitrev(F[2]), bitrev(F[3])};
    uint32_t S = 0;
#pragma unroll
    for (int i = 0; i < 28; i++)
        if (i < 14 || msgbits == 112) S ^= tab.nib_syn[(i + (off >> 2)) * 16 + frame_nibble(W, i)];
    errorbit = 0xFF; nfixed = 0;
    if (S != 0 && fix_errors && (msgtype == 11 || msgtype == 17 || msgtype == 18)) {
        const int pmin = off > 5 ? off : 5;             // table positions 5..111 (dump1090.c:806)
        const int hit = syndrome_pos(tab.fix_hash, S);
        if (hit >= pmin) {
            flip_bit(F, hit - off);
            errorbit = (uint32_t)(hit - off); nfixed = 1; S = 0;
        } else if (aggressive) {
            // two flipped bits p < q (all 5671 patterns have distinct syndromes): p from the pair
            // table, q as the position whose syndrome is S ^ syn[p]; both must lie in the frame
            const int p = pair_first(tab.pair_hash, S);
            if (p >= pmin) {
                const int q = syndrome_pos(tab.fix_hash, S ^ tab.bit_syn[p]);
                flip_bit(F, p - off); flip_bit(F, q - off);
                errorbit = (uint32_t)(p - off); nfixed = 2; S = 0;
            }
        }
    }
    crc = S;
}

// One attempt after slicing: message type, delta gate (dump1090.c:1713-1726), CRC and repair.
MODES_SERIAL_FN void judge_sliced(const uint32_t Fs[4], uint32_t tri, uint32_t sum56, uint32_t sum112, int fix_errors,
                                  int aggressive, const Tables &tab, Verdict &R) {
    R.F[0] = Fs[0]; R.F[1] = Fs[1]; R.F[2] = Fs[2]; R.F[3] = Fs[3];
    R.msgtype = bitrev(Fs[0]) >> 27;
    const int msgbits = (R.msgtype >= 16 && R.msgtype <= 21) ? 112 : 56;       // dump1090.c:746-753
    const uint32_t delta = (msgbits == 112) ? sum112 / 56u : sum56 / 28u;       // :1713-1718
    R.flags = tri ? MODES_EVAL_ERRORS : 0;
    R.crc = 0; R.errorbit = 0xFF; R.nfixed = 0;
    if (delta >= 2550u) {                                                       // :1723
        R.flags |= MODES_EVAL_GATE_OK;
        if (!tri || aggressive) {                                               // :1731 (errors is 0 or 1)
            R.flags |= MODES_EVAL_DECODED;
            crc_and_fix(R.F, msgbits, R.msgtype, fix_errors, aggressive, tab, R.crc, R.errorbit, R.nfixed);
        }
    }
}

MODES_SERIAL_FN bool unconditionally_good(const Verdict &R) {
    return (R.flags & MODES_EVAL_DECODED) && R.crc == 0 && (R.msgtype == 11 || R.msgtype == 17 || R.msgtype == 18);
}

// The six 32-bit words of a modes_frame_eval as laid out in memory.
MODES_SERIAL_FN void eval_words(const Verdict &R, uint32_t w[6]) {
    const uint32_t W0 = bitrev(R.F[0]), W1 = bitrev(R.F[1]), W2 = bitrev(R.F[2]), W3 = bitrev(R.F[3]) & 0xffff0000u;
    w[0] = bswap(W0); w[1] = bswap(W1); w[2] = bswap(W2);              // big-endian words -> bytes in memory order
    w[3] = (W3 >> 24) | ((W3 >> 8) & 0xff00u) | ((R.msgtype & 0xffu) << 16) | ((R.flags & 0xffu) << 24);
    w[4] = (R.errorbit & 0xffu) | ((R.nfixed & 0xffu) << 8);
    w[5] = R.crc & 0xffffffu;
}

MODES_SERIAL_FN uint32_t scale_sample(uint32_t v, uint32_t s) {               // dump1090.c:1473-1476
    const uint32_t r = (v * s) >> 14;
    return r > 65535u ? 65535u : r;
}

// "bits[0] == 2" (dump1090.c:1681): the first pair is a tie.  The 2 is copied into every
// following indefinite bit (:1675); packed with << (7 - k%8) (:1696-1706) a 2 at bit k sets bit
// k-1 unless k starts a byte, and leaves bit k clear.  `pairs` holds the magnitude pairs.
MODES_SERIAL_FN void spread_tie(const uint32_t *pairs, uint32_t F[4]) {
    for (int k = 1; k < 112; k++) {
        const uint32_t m = pairs[k];
        int d = (int)(m & 0xffffu) - (int)(m >> 16);
        d = d < 0 ? -d : d;
        if (d >= 256) break;                             // definite: the run ends
        if (k & 7) { const int b = k - 1; const uint32_t bit = 1u << (b & 31); const int w = b >> 5;
                     F[0] |= (w == 0) ? bit : 0u; F[1] |= (w == 1) ? bit : 0u; F[2] |= (w == 2) ? bit : 0u; F[3] |= (w == 3) ? bit : 0u; }
    }
}

// OR sixteen frame bits (bits 16k .. 16k+15) into F without dynamic register indexing.
MODES_SERIAL_FN void put16(uint32_t F[4], int k, uint32_t f16) {
    const uint32_t v = f16 << ((k & 1) * 16);
    const int w = k >> 1;
    F[0] |= (w == 0) ? v : 0u; F[1] |= (w == 1) ? v : 0u;
    F[2] |= (w == 2) ? v : 0u; F[3] |= (w == 3) ? v : 0u;
}

// The three per-bit loops below handle the 112 bits as 7 blocks of 16 (the block body is
// unrolled, the block loop is not: ~50 instructions per block keeps the kernel small).

// First pass, block k: raw samples -> magnitude pairs (stored back into the window) -> sliced
// bits (dump1090.c:1667-1690); accumulates the delta sums (:1692-1693).
MODES_SERIAL_FN uint32_t first_pass_block(uint32_t *win, int k, uint32_t shift, const uint16_t *lut_iq, uint32_t &wa,
                                          uint32_t &prev, uint32_t &dsum, uint32_t &d56) {
    uint32_t f = 0;
    uint32_t *p = win + 8 + 16 * k;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t wb = p[i + 1];
        const uint32_t a = absdiff127x4(funnel(wa, wb, shift));
        wa = wb;
        const int lo = (int)lut_iq[dot4(a, kLutLow)];
        const int hi = (int)lut_iq[dot4(a, kLutHigh)];
        p[i] = (uint32_t)lo | ((uint32_t)hi << 16);
        int d = lo - hi; d = d < 0 ? -d : d;
        dsum += (uint32_t)d;
        if (i == 7 && k == 3) d56 = dsum;                                      // bits 0..55 (:1716)
        bool definite = d >= 256;
        if (i == 0) definite = definite || k == 0;                             // bit 0 is always taken (:1675)
        prev = definite ? (uint32_t)(lo > hi) : prev;
        f |= prev << i;
    }
    return f;
}

// Phase correction (dump1090.c:1498-1558) and slicing of the corrected samples, one block of 16
// bits walking from *p in direction `step` (+1 forwards along the frame, -1 backwards): the
// half-bit sample next to the previous decision is rescaled by f_one or f_zero according to that
// decision, then the corrected pair is classified like the first pass (:1675): `d16` collects the
// definite bits, `o16` the definite ones, both in walk order.  `end0` / `end15` say that the
// first / last bit of the block is bit 0 of the frame (always definite; a tie there is the
// tri-state of :1681).
MODES_SERIAL_FN void correct_slice_block(const uint32_t *p, int step, bool fwd, bool end0, bool end15, uint32_t f_one,
                                         uint32_t f_zero, uint32_t &prev_e, uint32_t &d16, uint32_t &o16, uint32_t &tie0) {
    d16 = 0; o16 = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t m = p[i * step];
        const uint32_t lo = m & 0xffffu, hi = m >> 16;
        const uint32_t xs = scale_sample(fwd ? lo : hi, prev_e ? f_one : f_zero);
        const int nlo = (int)(fwd ? xs : lo), nhi = (int)(fwd ? hi : xs);
        prev_e = (uint32_t)(nlo > nhi);
        int d = nlo - nhi; d = d < 0 ? -d : d;
        bool definite = d >= 256;
        if (i == 0 && end0) { definite = true; tie0 = (uint32_t)(nlo == nhi); }
        if (i == 15 && end15) { definite = true; tie0 = (uint32_t)(nlo == nhi); }
        d16 |= (uint32_t)definite << i;
        o16 |= (uint32_t)(definite && nlo > nhi) << i;
    }
}

// Leaner code for the same first-pass block (|lo - hi| as one VABSDIFF; 18.7 instead of 24.6
// instructions per bit).  First pass, block k: raw samples -> magnitude pairs (stored back into the window) -> sliced
// bits (dump1090.c:1667-1690); accumulates the delta sums (:1692-1693).
MODES_SERIAL_FN uint32_t first_pass_block_lean(uint32_t *win, int k, uint32_t shift, const uint16_t *lut_iq, uint32_t &wa,
                                          uint32_t &prev, uint32_t &dsum, uint32_t &d56) {
    uint32_t f = 0;
    uint32_t *p = win + 8 + 16 * k;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t wb = p[i + 1];
        const uint32_t a = absdiff127x4(funnel(wa, wb, shift));
        wa = wb;
        const uint32_t lo = lut_iq[dot4(a, kLutLow)];
        const uint32_t hi = lut_iq[dot4(a, kLutHigh)];
        p[i] = lo | (hi << 16);
        const uint32_t d = absdiff_add(lo, hi, 0u);
        dsum += d;
        if (i == 7 && k == 3) d56 = dsum;                                      // bits 0..55 (:1716)
        bool definite = d >= 256u;
        if (i == 0) definite = definite || k == 0;                             // bit 0 is always taken (:1675)
        prev = definite ? (uint32_t)(lo > hi) : prev;
        f |= prev << i;
    }
    return f;
}

// Leaner code for the same block (20.3 instead of 24.8 instructions per bit).
// Phase correction (dump1090.c:1498-1558) and slicing of the corrected samples, one block of 16
// bits walking from *p in direction `step` (+1 forwards along the frame, -1 backwards): the
// half-bit sample next to the previous decision is rescaled by f_one or f_zero according to that
// decision, then the corrected pair is classified like the first pass (:1675): `d16` collects the
// definite bits, `o16` the definite ones, both in walk order.  `end0` / `end15` say that the
// first / last bit of the block is bit 0 of the frame (always definite; a tie there is the
// tri-state of :1681).
//
// Written on integer masks rather than bools (all-ones = true): with sixteen unrolled steps the
// compiler otherwise shuttles the conditions between predicates and registers (7 SEL per bit).
//   sel_x / sel_y  byte selectors that extract the rescaled / the other half of a pair, zero-extended
//   dir            0 forwards, ~0 backwards;   f_pick = f_one ^ f_zero
//   not_e          ~0 unless the previous decision was "one"
MODES_SERIAL_FN void correct_slice_block_lean(const uint32_t *p, int step, uint32_t sel_x, uint32_t sel_y, uint32_t dir,
                                         bool end0, bool end15, uint32_t f_one, uint32_t f_pick, uint32_t &not_e,
                                         uint32_t &d16, uint32_t &o16, uint32_t &tie0) {
    uint32_t nd = 0, no = 0;                               // complements of d16 / o16, built by OR
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const uint32_t m = p[i * step];
        const uint32_t x = perm(m, sel_x), y = perm(m, sel_y);
        const uint32_t xs = scale_sample(x, f_one ^ (not_e & f_pick));          // previous one: f_one, else f_zero
        const int32_t s = (int32_t)(xs - y);
        // "one" = first corrected sample > second: forwards s > 0, backwards s < 0.
        // u = forwards s-1, backwards ~s: one <=> u >= 0
        const int32_t u = (int32_t)(((uint32_t)s ^ dir) + ~dir);
        not_e = (uint32_t)(u >> 31);
        uint32_t not_def = (uint32_t)((int32_t)(absdiff_add(xs, y, 0u) - 256u) >> 31);   // ~0 when |delta| < 256
        if (i == 0 && end0) { not_def = 0; tie0 = (uint32_t)(s == 0); }
        if (i == 15 && end15) { not_def = 0; tie0 = (uint32_t)(s == 0); }
        nd |= not_def & (1u << i);
        no |= (not_def | not_e) & (1u << i);
    }
    d16 = ~nd & 0xffffu;
    o16 = ~no & 0xffffu;
}

// Bits with D=1 are definite and take their value from O; bits with D=0 copy the nearest
// definite bit below (:1675).  That is the carry into each bit of (O|~D) + O.  Bit 0 must be in D.
MODES_SERIAL_FN void fill_copies(const uint32_t D[4], const uint32_t O[4], uint32_t F[4]) {
    const uint64_t d0 = D[0] | ((uint64_t)D[1] << 32), d1 = D[2] | ((uint64_t)D[3] << 32);
    const uint64_t o0 = O[0] | ((uint64_t)O[1] << 32), o1 = O[2] | ((uint64_t)O[3] << 32);
    const uint64_t a0 = o0 | ~d0, a1 = o1 | ~d1;
    const uint64_t s0 = a0 + o0;
    const uint64_t s1 = a1 + o1 + (s0 < a0 ? 1ull : 0ull);
    const uint64_t ci0 = s0 ^ a0 ^ o0, ci1 = s1 ^ a1 ^ o1;                   // carry INTO each bit
    const uint64_t f0 = (ci0 >> 1) | (ci1 << 63), f1 = (ci1 >> 1) & 0x0000ffffffffffffull;
    F[0] = (uint32_t)f0; F[1] = (uint32_t)(f0 >> 32); F[2] = (uint32_t)f1; F[3] = (uint32_t)(f1 >> 32);
}

// spread_tie on the definite mask instead of the pairs: the run of indefinite bits after bit 0.
MODES_SERIAL_FN void spread_tie_mask(const uint32_t D[4], uint32_t F[4]) {
    const uint64_t t0 = ~(D[0] | ((uint64_t)D[1] << 32)) | 1ull, t1 = ~(D[2] | ((uint64_t)D[3] << 32));
    const uint64_t u0 = t0 + 1ull, u1 = t1 + (u0 == 0 ? 1ull : 0ull);
    const uint64_t run0 = t0 & ~u0, run1 = (u0 == 0) ? (t1 & ~u1) : 0ull;       // bits 0..r
    const uint64_t e0 = ((run0 >> 1) | (run1 << 63)) & 0x7f7f7f7f7f7f7f7full;
    const uint64_t e1 = (run1 >> 1) & 0x00007f7f7f7f7f7full;
    F[0] |= (uint32_t)e0; F[1] |= (uint32_t)(e0 >> 32); F[2] |= (uint32_t)e1; F[3] |= (uint32_t)(e1 >> 32);
}

MODES_SERIAL_FN uint32_t window_sample(const uint32_t *win, uint32_t odd, int w) {
    const uint32_t h = (uint32_t)w + odd;
    const uint32_t word = win[h >> 1];
    return (h & 1u) ? (word >> 16) : (word & 0xffffu);
}

MODES_SERIAL_FN uint32_t magnitude_of(const uint16_t *lut_iq, uint32_t iq) {
    return lut_iq[dot4(absdiff127x4(iq | 0x7f7f0000u), kLutLow)];
}

// Evaluate one candidate.  win: its kWindowWords staged words (modified); odd: see above;
// at_buffer_start: j == 0, where the reference retries without correcting (dump1090.c:1660);
// rec: the 12 words of pass[0] and pass[1] of its modes_candidate.
template <bool kLean>
MODES_SERIAL_FN void evaluate(uint32_t *win, uint32_t odd, bool at_buffer_start, int fix_errors, int aggressive,
                              const Tables &tab, uint32_t rec[12]) {
    const uint32_t shift = odd ? 32u : 16u;
    uint32_t F[4] = {0u, 0u, 0u, 0u};
    uint32_t prev = 0, dsum = 0, d56 = 0;
    uint32_t wa = win[8];
#pragma unroll 1
    for (int k = 0; k < 7; k++) {
        if constexpr (kLean) put16(F, k, first_pass_block_lean(win, k, shift, tab.lut_iq, wa, prev, dsum, d56));
        else put16(F, k, first_pass_block(win, k, shift, tab.lut_iq, wa, prev, dsum, d56));
    }
    const uint32_t d112 = dsum;
    uint32_t *pairs = win + 8;
    const uint32_t tri1 = (pairs[0] & 0xffffu) == (pairs[0] >> 16);
    if (tri1) spread_tie(pairs, F);

    Verdict P1, P2;
    judge_sliced(F, tri1, d56, d112, fix_errors, aggressive, tab, P1);
    P2.F[0] = P2.F[1] = P2.F[2] = P2.F[3] = 0;
    P2.msgtype = 0; P2.flags = 0; P2.errorbit = 0; P2.nfixed = 0; P2.crc = 0;

    if ((P1.flags & MODES_EVAL_GATE_OK) && !unconditionally_good(P1)) {
        P1.flags |= MODES_EVAL_P2_VALID;
        if (at_buffer_start) {
            P2 = P1;
            P2.flags &= ~(uint32_t)MODES_EVAL_P2_VALID;
        } else {
            // applyPhaseCorrection, dump1090.c:1498-1558
            const uint32_t m_1 = magnitude_of(tab.lut_iq, window_sample(win, odd, 0));
            const uint32_t m0 = magnitude_of(tab.lut_iq, window_sample(win, odd, 1));
            const uint32_t m2 = magnitude_of(tab.lut_iq, window_sample(win, odd, 3));
            const uint32_t m3 = magnitude_of(tab.lut_iq, window_sample(win, odd, 4));
            const uint32_t m6 = magnitude_of(tab.lut_iq, window_sample(win, odd, 7));
            const uint32_t m7 = magnitude_of(tab.lut_iq, window_sample(win, odd, 8));
            const uint32_t m9 = magnitude_of(tab.lut_iq, window_sample(win, odd, 10));
            const uint32_t m10 = magnitude_of(tab.lut_iq, window_sample(win, odd, 11));
            const uint32_t on_time = m0 + m2 + m7 + m9;
            const uint32_t early = (m_1 + m6) * 2u, late = (m3 + m10) * 2u;
            // early > late: walk backwards, the second half-bit samples are rescaled;
            // otherwise forwards, the first half-bit samples are.
            const bool fwd = !(early > late);
            const uint32_t lead = fwd ? late : early;
            const uint32_t q = 16384u * lead / (lead + on_time);
            const uint32_t up = (16384u + q) & 0xffffu, down = (16384u - q) & 0xffffu;
            // forwards a previous 1 scales up, backwards a following 1 scales down; the first
            // bit handled always scales up
            const uint32_t f_one = fwd ? up : down, f_zero = fwd ? down : up;
            uint32_t Dm[4] = {0u, 0u, 0u, 0u}, Om[4] = {0u, 0u, 0u, 0u}, tri2 = 0;
            if constexpr (kLean) {
                const int step = fwd ? 1 : -1;
                // forwards the first half-bit sample (low half of a pair) is rescaled, backwards the second
                const uint32_t sel_x = fwd ? 0x4410u : 0x4432u, sel_y = fwd ? 0x4432u : 0x4410u;
                const uint32_t dir = fwd ? 0u : ~0u;
                const uint32_t g_one = opaque(f_one), g_pick = opaque(f_one ^ f_zero);
                uint32_t not_e = fwd ? 0u : ~0u;               // the first bit handled always scales up
                const uint32_t *cp = fwd ? pairs : pairs + 111;
    #pragma unroll 1
                for (int k = 0; k < 7; k++, cp += 16 * step) {
                    uint32_t d16, o16;
                    correct_slice_block_lean(cp, step, sel_x, sel_y, dir, fwd && k == 0, !fwd && k == 6, g_one, g_pick, not_e, d16, o16, tri2);
                    if (!fwd) { d16 = bitrev(d16) >> 16; o16 = bitrev(o16) >> 16; }      // walk order -> frame order
                    put16(Dm, fwd ? k : 6 - k, d16);
                    put16(Om, fwd ? k : 6 - k, o16);
                }
            } else {
                uint32_t prev_e = fwd ? 1u : 0u;
                const int step = fwd ? 1 : -1;
                const uint32_t *cp = fwd ? pairs : pairs + 111;
    #pragma unroll 1
                for (int k = 0; k < 7; k++, cp += 16 * step) {
                    uint32_t d16, o16;
                    correct_slice_block(cp, step, fwd, fwd && k == 0, !fwd && k == 6, f_one, f_zero, prev_e, d16, o16, tri2);
                    if (!fwd) { d16 = bitrev(d16) >> 16; o16 = bitrev(o16) >> 16; }      // walk order -> frame order
                    put16(Dm, fwd ? k : 6 - k, d16);
                    put16(Om, fwd ? k : 6 - k, o16);
                }
            }
            uint32_t G[4];
            fill_copies(Dm, Om, G);
            if (tri2) spread_tie_mask(Dm, G);
            if (G[0] == F[0] && G[1] == F[1] && G[2] == F[2] && G[3] == F[3] && tri2 == tri1) {
                P2 = P1;                                 // same bits, same (uncorrected) delta sums: same verdict
                P2.flags &= ~(uint32_t)MODES_EVAL_P2_VALID;
            } else {
                judge_sliced(G, tri2, d56, d112, fix_errors, aggressive, tab, P2);
            }
        }
    }
    eval_words(P1, rec);
    eval_words(P2, rec + 6);
}

}  // namespace serial
}  // namespace modes
