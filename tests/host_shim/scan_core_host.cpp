// TEST INFRASTRUCTURE — not part of the product.
// Host build of dump1090_b200/csrc/modes_scan_core.cuh (one lane's share of one row of the
// preamble scan kernel), so that the packed arithmetic can be checked in the CPU test suite
// against a direct evaluation of the ten comparisons of dump1090.c:1602-1611.
#include <cstring>
#include "modes_scan_core.cuh"

// iq: n_samples I/Q pairs.  mask_out: one bit per position p in [0, n_positions), computed in
// groups of 32 positions exactly as a lane does (16 own words + 5 lookahead words); the caller
// supplies at least n_positions + 10 samples.
extern "C" int shim_scan_masks(const uint8_t *iq, uint64_t n_samples, uint64_t n_positions, uint32_t *mask_out) {
    using namespace modes::scan2;
    if (n_positions % 32 || n_samples < n_positions + 2 * kLookWords) return -1;
    for (uint64_t p0 = 0; p0 < n_positions; p0 += 32) {
        uint32_t P[kLaneWords + kLookWords];
        for (int k = 0; k < kLaneWords + kLookWords; k++) {
            uint32_t raw;
            std::memcpy(&raw, iq + 2 * (p0 + 2 * k), 4);
            P[k] = npack(raw);
        }
        mask_out[p0 / 32] = row_mask(P, 1u, 0xffffffffu);
    }
    return 0;
}

// Two rows at a time (rows2_mask): row A = positions [0, n_positions), row B starts row_b samples later.
extern "C" int shim_scan_masks2(const uint8_t *iq, uint64_t n_samples, uint64_t n_positions, uint64_t row_b,
                                uint32_t *mask_a, uint32_t *mask_b) {
    using namespace modes::scan2;
    if (n_positions % 32 || n_samples < row_b + n_positions + 2 * kPairLook + 2) return -1;
    for (uint64_t p0 = 0; p0 < n_positions; p0 += 32) {
        uint32_t X[32 + kPairLook + 1];
        for (int w = 0; w < (32 + kPairLook + 1) / 2; w++) {
            uint32_t ra, rb;
            std::memcpy(&ra, iq + 2 * (p0 + 2 * w), 4);
            std::memcpy(&rb, iq + 2 * (row_b + p0 + 2 * w), 4);
            npack2(ra, rb, 65536u, 0xffffffffu, X[2 * w], X[2 * w + 1]);
        }
        rows2_mask(X, 1u, 0xffffffffu, mask_a[p0 / 32], mask_b[p0 / 32]);
    }
    return 0;
}

extern "C" uint32_t shim_npack(uint32_t raw) { return modes::scan2::npack(raw); }
