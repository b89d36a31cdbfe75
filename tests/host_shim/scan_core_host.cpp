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

extern "C" uint32_t shim_npack(uint32_t raw) { return modes::scan2::npack(raw); }
