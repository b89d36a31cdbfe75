// modes_scan2.cu — K1, the fused magnitude + preamble scan kernel (sm_100a).
//
// Replaces computeMagnitudeVector (dump1090.c:1454-1469) and the per-position tests of
// detectModeS (dump1090.c:1602-1650).  HBM-bound by design: 2 bytes read per sample, nothing
// written but the sparse candidate list; in practice the limit is the integer instruction rate,
// so the kernel is built around instructions per sample (modes_scan_core.cuh has the arithmetic).
//
// Work split.  One warp owns a tile of 4096 positions and walks it in 4 rows of 1024: in row r
// lane l handles the 32 consecutive positions 1024r + 32l .. +31 (64 bytes of I/Q = four 16-byte
// loads; the warp's loads of a row cover 2 KB contiguously).  The 10 samples of lookahead the
// last positions of a lane need are the first five packed words of the next lane, fetched by
// shuffle (lane 31 takes them from the next row, whose words are packed one row early).  A
// lane's 32 pass flags come out as one register, so no shared-memory transposition is needed.
// Rows are requested two ahead of their use, the next tile's first rows before the current tile's
// survivor stage; tiles are handed out from a global counter, requested one tile ahead.
//
// Survivors of the ten comparisons (~1 % of positions) get the exact "high" tests of
// dump1090.c:1624-1642 on table magnitudes, 32 at a time in position order; their samples are
// re-read through L1, where the row loads have just put them.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "modes_internal.h"
#include "modes_scan_core.cuh"

namespace modes {
namespace {

constexpr int kRows = 4;                                 // rows per tile
constexpr int kRowChunks = 128;                          // 16-byte chunks (8 samples) per row
constexpr int kTileChunks = kTileSamples / 8;
constexpr int kSurvivorCap = 512;
constexpr int kOutCap = 256;                             // candidates per tile held back one tile
constexpr int kSmemBytes = kSurvivorCap * 2 + 2 * kOutCap * 2;
static_assert(kTileSamples == kRows * 1024, "tile = 4 rows of 32 lanes x 32 positions");

struct TileSrc {
    const uint4 *flat;           // chunk 0 of the tile when the tile (+ lookahead) lies inside the body
    uint64_t c0;                 // first virtual chunk of the tile
    bool interior;
};

__device__ __forceinline__ TileSrc tile_source(const BatchView &in, uint32_t g, uint64_t n_vchunks) {
    TileSrc ts;
    ts.c0 = (uint64_t)g * kTileChunks;
    ts.interior = ts.c0 >= kHaloSamples / 8 && ts.c0 + kTileChunks + 3 <= n_vchunks;
    ts.flat = reinterpret_cast<const uint4 *>(in.body) + (ts.c0 - kHaloSamples / 8);
    return ts;
}

// Chunk c of the virtual sample array; chunks past the end read as "no signal" (127,127).
__device__ __forceinline__ uint4 load_vchunk(const BatchView &in, uint64_t c, uint64_t n_vchunks) {
    if (c >= n_vchunks) return make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
    const uint4 *p = (c < kHaloSamples / 8) ? reinterpret_cast<const uint4 *>(in.halo) + c
                                            : reinterpret_cast<const uint4 *>(in.body) + (c - kHaloSamples / 8);
    return __ldg(p);
}

// The rows are loaded through L1 (no L1::no_allocate): the exact tests re-read a few of their samples.
__device__ __forceinline__ uint4 load_chunk(const BatchView &in, const TileSrc &t, int chunk, uint64_t n_vchunks) {
    if (t.interior) return __ldg(t.flat + chunk);
    return load_vchunk(in, t.c0 + chunk, n_vchunks);
}

__device__ __forceinline__ void load_row(const BatchView &in, const TileSrc &t, int r, int lane, uint64_t n_vchunks, uint4 x[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) x[p] = load_chunk(in, t, kRowChunks * r + 4 * lane + p, n_vchunks);
}

__device__ __forceinline__ void pack_row(const uint4 x[4], uint32_t P[16]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
        P[4 * p + 0] = scan2::npack(x[p].x); P[4 * p + 1] = scan2::npack(x[p].y);
        P[4 * p + 2] = scan2::npack(x[p].z); P[4 * p + 3] = scan2::npack(x[p].w);
    }
}

// Pass flags of this lane's 32 positions in the row packed in Pc; Pn = the next row's first words
// (lane 31's lookahead is lane 0's share of the next row).
__device__ __forceinline__ uint32_t scan_row(const uint32_t Pc[16], const uint32_t Pn[scan2::kLookWords], int lane,
                                             uint32_t one, uint32_t minus_one) {
    uint32_t P[scan2::kLaneWords + scan2::kLookWords];
#pragma unroll
    for (int k = 0; k < 16; k++) P[k] = Pc[k];
#pragma unroll
    for (int k = 0; k < scan2::kLookWords; k++)
        P[16 + k] = __shfl_sync(0xffffffffu, lane == 0 ? Pn[k] : Pc[k], (lane + 1) & 31);
    return scan2::row_mask(P, one, minus_one);
}

// Squared magnitude of tile sample s (0 .. 4096+14), re-read from global memory (L1 / L2).
__device__ __forceinline__ uint32_t tile_n(const BatchView &in, const TileSrc &t, int s) {
    uint32_t w;
    if (t.interior) w = __ldg(reinterpret_cast<const uint16_t *>(t.flat) + s);
    else {
        const uint64_t v = t.c0 * 8 + (uint64_t)s;
        const uint8_t *p = (v < (uint64_t)kHaloSamples) ? in.halo + 2 * v : in.body + 2 * (v - kHaloSamples);
        w = *reinterpret_cast<const uint16_t *>(p);
    }
    const uint32_t a = __vabsdiffu4(w | 0x7f7f0000u, 0x7f7f7f7fu);
    return __dp4a(a, a, 0u);
}

// dump1090.c:1624-1642 for tile-local position s, on exact magnitudes:
//   high = (m0+m2+m7+m9)/6;  m4, m5, m11..m14 < high
// <=> 6*(max(m4,m5,m11..m14)+1) <= m0+m2+m7+m9, and the magnitude table is monotone in the
// squared magnitude, so the max is taken before the lookup: five lookups instead of ten.
__device__ __forceinline__ bool high_tests(const BatchView &in, const TileSrc &t, int s, const uint16_t *__restrict__ lutn) {
#define TN(d) tile_n(in, t, s + (d))
    const uint32_t n0 = TN(0), n2 = TN(2), n7 = TN(7), n9 = TN(9);
    const uint32_t n4 = TN(4), n5 = TN(5), n11 = TN(11), n12 = TN(12), n13 = TN(13), n14 = TN(14);
#undef TN
    const uint32_t nx = max(max(max(n4, n5), max(n11, n12)), max(n13, n14));
    const int sum = (int)__ldg(lutn + n0) + (int)__ldg(lutn + n2) + (int)__ldg(lutn + n7) + (int)__ldg(lutn + n9);
    const int mx = __ldg(lutn + nx);
    return 6 * (mx + 1) <= sum;
}

// This lane's survivors (bit i of m[r] = tile position 1024r + 32*lane + i) -> the slots
// [excl[r], ...) of the tile-ordered survivor sequence; those in [round, round+kSurvivorCap) are listed.
__device__ __forceinline__ void list_survivors(const uint32_t m[kRows], const uint32_t excl[kRows], uint32_t round,
                                               uint16_t *surv, int lane) {
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        uint32_t slot = excl[r] - round;
        for (uint32_t b = m[r]; b; b &= b - 1) {
            if (slot < (uint32_t)kSurvivorCap) surv[slot] = (uint16_t)(1024 * r + 32 * lane + __ffs(b) - 1);
            slot++;
        }
    }
}

// Copy a finished tile's candidate list (tile-local positions, position order) to its slot in the
// global candidate array and record the tile.  `base0` is the slot claimed one tile earlier.
__device__ __forceinline__ void emit_tile(const ScanOutputs &out, const uint16_t *olist, uint32_t tile, uint32_t base0,
                                          uint32_t total, int lane) {
    const uint32_t base = __shfl_sync(0xffffffffu, base0, 0);
    const uint32_t v0 = tile * (uint32_t)kTileSamples;
    for (uint32_t i = lane; i < total; i += 32)
        if (base + i < out.cand_capacity) out.cand_v[base + i] = v0 + olist[i];
    if (lane == 0) {
        uint32_t stored = total;
        if (base + total > out.cand_capacity) {
            stored = base < out.cand_capacity ? out.cand_capacity - base : 0;
            out.counters[1] = 1;
        }
        modes_tile tl; tl.offset = base; tl.count = stored;
        out.tiles[tile] = tl;
    }
}

__global__ void __launch_bounds__(32, 16)
scan2_kernel(BatchView in, const uint16_t *__restrict__ lutn, ScanOutputs out, uint32_t n_tiles, uint32_t one,
             uint32_t minus_one) {
    __shared__ __align__(16) uint16_t s_surv[kSurvivorCap];
    __shared__ __align__(16) uint16_t s_olist[2 * kOutCap];
    uint32_t pend_tile = 0xffffffffu, pend_base = 0, pend_total = 0, pend_buf = 0;

    const int lane = threadIdx.x;
    const uint64_t n_vchunks = (in.n_samples + kHaloSamples) / 8;
    const uint64_t t_end = in.n_samples;

    uint32_t g = blockIdx.x;
    if (g >= n_tiles) return;
    TileSrc ts = tile_source(in, g, n_vchunks);
    uint4 xa[4], xb[4];                                   // rows in flight: even / odd
    load_row(in, ts, 0, lane, n_vchunks, xa);
    load_row(in, ts, 1, lane, n_vchunks, xb);

    // Tiles are handed out first come, first served (after one static tile per warp); the index
    // is requested a whole tile before it is needed, so the atomic's latency is never waited for.
    uint32_t g_ahead = 0;
    if (lane == 0) g_ahead = gridDim.x + atomicAdd(&out.counters[2], 1u);

    for (int it = 0; g < n_tiles; ++it) {
        const int cur = it & 1;
        const TileSrc ts_cur = ts;
        uint32_t m[kRows];
        uint32_t Pc[16], Pn[16];
        pack_row(xa, Pc);

        // row 0: request row 2, pack row 1
        load_row(in, ts_cur, 2, lane, n_vchunks, xa);
        pack_row(xb, Pn);
        m[0] = scan_row(Pc, Pn, lane, one, minus_one);
#pragma unroll
        for (int k = 0; k < 16; k++) Pc[k] = Pn[k];

        // row 1: request row 3, pack row 2
        load_row(in, ts_cur, 3, lane, n_vchunks, xb);
        pack_row(xa, Pn);
        m[1] = scan_row(Pc, Pn, lane, one, minus_one);
#pragma unroll
        for (int k = 0; k < 16; k++) Pc[k] = Pn[k];

        // row 2: request the start of the row after the tile (lane 31's lookahead in row 3; lane 0
        // fetches it) and the next tile's row 0; pack row 3
        const uint4 pad = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
        uint4 xe0 = pad, xe1 = pad;
        if (lane == 0) {
            xe0 = load_chunk(in, ts_cur, kTileChunks, n_vchunks);
            xe1 = load_chunk(in, ts_cur, kTileChunks + 1, n_vchunks);
        }
        const uint32_t g_next = __shfl_sync(0xffffffffu, g_ahead, 0);
        if (lane == 0 && g_next < n_tiles) g_ahead = gridDim.x + atomicAdd(&out.counters[2], 1u);
        if (g_next < n_tiles) {
            ts = tile_source(in, g_next, n_vchunks);
            load_row(in, ts, 0, lane, n_vchunks, xa);
        }
        pack_row(xb, Pn);
        m[2] = scan_row(Pc, Pn, lane, one, minus_one);
#pragma unroll
        for (int k = 0; k < 16; k++) Pc[k] = Pn[k];

        // row 3: request the next tile's row 1
        if (g_next < n_tiles) load_row(in, ts, 1, lane, n_vchunks, xb);
        {
            uint32_t Pe[scan2::kLookWords];
            Pe[0] = scan2::npack(xe0.x); Pe[1] = scan2::npack(xe0.y); Pe[2] = scan2::npack(xe0.z);
            Pe[3] = scan2::npack(xe0.w); Pe[4] = scan2::npack(xe1.x);
            m[3] = scan_row(Pc, Pe, lane, one, minus_one);
        }

        // ---- write out the PREVIOUS tile's candidates: its slot in the global array (one atomic
        // per tile) was claimed before this tile's rows were scanned, so the round trip is hidden
        if (pend_tile != 0xffffffffu) {
            emit_tile(out, s_olist + kOutCap * pend_buf, pend_tile, pend_base, pend_total, lane);
            pend_tile = 0xffffffffu;
        }

        // ---- positions the reference never tests (dump1090.c:1593): j >= 131070 are the first two
        // positions of every 32nd tile (v = t+2); the last tile ends at t = N-1
        const uint32_t v_tile = g * (uint32_t)kTileSamples;
        if ((g & 31u) == 0 && lane == 0) m[0] &= ~3u;
        if (t_end + 2 - v_tile < (uint64_t)kTileSamples) {
            const int s_max = (int)(t_end + 2 - v_tile);
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                const int lo_pos = 1024 * r + 32 * lane;
                if (lo_pos >= s_max) m[r] = 0;
                else if (lo_pos + 32 > s_max) m[r] &= (1u << (s_max - lo_pos)) - 1u;
            }
        }

        // ---- survivors in position order: row-major, lane, bit.  Two packed prefix sums.
        uint32_t c[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) c[r] = __popc(m[r]);
        uint32_t x = c[0] | (c[1] << 16), y = c[2] | (c[3] << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ox = __shfl_up_sync(0xffffffffu, x, d), oy = __shfl_up_sync(0xffffffffu, y, d);
            if (lane >= d) { x += ox; y += oy; }
        }
        const uint32_t tx = __shfl_sync(0xffffffffu, x, 31), ty = __shfl_sync(0xffffffffu, y, 31);
        uint32_t excl[kRows];
        {
            const uint32_t b1 = tx & 0xffffu, b2 = b1 + (tx >> 16), b3 = b2 + (ty & 0xffffu);
            excl[0] = (x & 0xffffu) - c[0];
            excl[1] = b1 + (x >> 16) - c[1];
            excl[2] = b2 + (y & 0xffffu) - c[2];
            excl[3] = b3 + (y >> 16) - c[3];
        }
        const uint32_t n_surv = (tx & 0xffffu) + (tx >> 16) + (ty & 0xffffu) + (ty >> 16);

        uint16_t *olist = s_olist + kOutCap * cur;
        uint32_t n_out = 0;
        bool dense = n_surv > (uint32_t)kSurvivorCap;
        if (!dense) {
            list_survivors(m, excl, 0, s_surv, lane);
            __syncwarp();
            for (uint32_t i0 = 0; i0 < n_surv; i0 += 32) {
                const uint32_t i = i0 + lane;
                const int spos = i < n_surv ? s_surv[i] : 0;
                const bool pass = i < n_surv && high_tests(in, ts_cur, spos, lutn);
                const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                const uint32_t slot = n_out + __popc(bal & ((1u << lane) - 1u));
                if (pass && slot < (uint32_t)kOutCap) olist[slot] = (uint16_t)spos;
                n_out += __popc(bal);
            }
            dense = n_out > (uint32_t)kOutCap;
        }
        if (!dense) {
            // claim the slot now, copy the list one tile later
            pend_base = 0;
            if (lane == 0 && n_out) pend_base = atomicAdd(&out.counters[0], n_out);
            pend_total = n_out; pend_tile = g; pend_buf = cur;
        } else {
            // pathological density: count, claim, then write straight to the global array
            uint32_t total = 0;
            for (int pass_no = 0; pass_no < 2; pass_no++) {
                uint32_t base = 0, run = 0;
                if (pass_no == 1) {
                    if (lane == 0 && total) base = atomicAdd(&out.counters[0], total);
                    base = __shfl_sync(0xffffffffu, base, 0);
                }
                for (uint32_t round = 0; round < n_surv; round += kSurvivorCap) {
                    __syncwarp();
                    list_survivors(m, excl, round, s_surv, lane);
                    __syncwarp();
                    const uint32_t n_here = min(n_surv - round, (uint32_t)kSurvivorCap);
                    for (uint32_t i0 = 0; i0 < n_here; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        const int spos = i < n_here ? s_surv[i] : 0;
                        const bool pass = i < n_here && high_tests(in, ts_cur, spos, lutn);
                        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
                        if (pass_no == 1 && pass) {
                            const uint32_t idx = base + run + __popc(bal & ((1u << lane) - 1u));
                            if (idx < out.cand_capacity) out.cand_v[idx] = v_tile + spos;
                        }
                        run += __popc(bal);
                    }
                }
                if (pass_no == 0) total = run;
                else if (lane == 0) {
                    uint32_t stored = total;
                    if (base + total > out.cand_capacity) {
                        stored = base < out.cand_capacity ? out.cand_capacity - base : 0;
                        out.counters[1] = 1;
                    }
                    modes_tile tl; tl.offset = base; tl.count = stored;
                    out.tiles[g] = tl;
                }
            }
        }
        __syncwarp();                                    // the lists are reused by the next tile
        g = g_next;
    }
    if (pend_tile != 0xffffffffu) emit_tile(out, s_olist + kOutCap * pend_buf, pend_tile, pend_base, pend_total, lane);
}

}  // namespace

void launch_scan2(const BatchView &in, const DeviceTables &tab, const ScanOutputs &out, int sm_count, cudaStream_t stream) {
    // occupancy is a property of the device the context lives on
    static int ctas_per_sm[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!ctas_per_sm[dev]) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan2_kernel, 32, 0) != cudaSuccess || n < 1) n = 16;
        ctas_per_sm[dev] = n;
    }
    const uint32_t n_tiles = tiles_for(in.n_samples);
    uint32_t grid = (uint32_t)(sm_count * ctas_per_sm[dev]);   // persistent single-warp CTAs, all resident
    if (grid > n_tiles) grid = n_tiles;
    scan2_kernel<<<grid, 32, 0, stream>>>(in, tab.lutn, out, n_tiles, 1u, 0xffffffffu);
}

}  // namespace modes
