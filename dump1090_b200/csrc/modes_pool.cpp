// modes_pool.cpp — many receivers on one GPU: "batch across receivers instead of across time"
// (SURVEY.md §8(f) item 4).
//
// One dump1090 process serves one RTL-SDR: rtlsdrCallback (dump1090.c:442-456) hands it 131072
// samples at a time, each buffer is prefixed with the last 238 samples of the previous one (:481),
// and detectModeS keeps its ICAO address cache, skip state and statistics for that one stream.  A
// B200 decodes ~12 000 such 2 MHz streams in real time (the limit is the PCIe link), but a single
// stream only fills it with seconds of data at a time.  The pool takes ONE buffer from each of many
// receivers and decodes them in one batch, with everything per stream kept per receiver.
//
// The kernels are untouched: they see a batch as consecutive buffers of ONE stream, each buffer's
// carry being the tail of the buffer before it.  The batch is therefore laid out as pairs
//     [pad buffer of receiver i: no signal, its last 238 samples = receiver i's carry][receiver i's buffer]
// resident in HBM; per call only the data buffers and the 476-byte carries cross PCIe.  A position
// of a data buffer never looks beyond that buffer and its 238 carried samples (dump1090.c:1593:
// j < 131070, window j..j+239 of a 131310-sample array), so its candidates are exactly those of the
// receiver's own stream; whatever the scan finds inside the pad buffers is dropped.  The price is a
// second scan over constant buffers (the scan runs at 2 T samples/s; the link delivers 25 G samples/s).
//
// The order-dependent half is the existing host resolver, one per receiver: the records of a data
// buffer are re-based to that receiver's own stream position and replayed with its own address
// cache.  Built on the public C ABI only (modes_detect_device / _fetch, modes_resolver_*).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "modes_b200.h"

namespace {

constexpr size_t kBuf = MODES_BUFFER_BYTES;
constexpr int64_t kBufSamples = MODES_BUFFER_SAMPLES;

struct Receiver {
    modes_resolver *res = nullptr;
    int64_t buffers = 0;                       // buffers of this receiver decoded so far
    uint8_t carry[MODES_CARRY_BYTES];
};

}  // namespace

struct modes_pool {
    modes_config cfg;
    std::vector<Receiver> rx;
    size_t max_batch = 0;
    std::string err;
    // device half (created on the first modes_pool_ingest)
    modes_ctx *ctx = nullptr;
    uint8_t *d_batch = nullptr;                // [2 * max_batch] buffers: pad, data, pad, data, ...
    uint8_t *h_carry = nullptr;                // pinned [max_batch][MODES_CARRY_BYTES]
    std::vector<modes_candidate> cands;
    std::vector<modes_tile> tiles;
    // scratch of the per-receiver resolve
    std::vector<modes_candidate> local;
    std::vector<modes_tile> local_tiles;
    // optional caller-owned output array (modes_pool_set_output)
    modes_message *out = nullptr; uint32_t *out_rx = nullptr; size_t out_cap = 0, out_count = 0;
};

namespace {

int fail(modes_pool *p, const char *fmt, ...) {
    char buf[256];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    p->err = buf;
    return -1;
}

int check_ids(modes_pool *p, const uint32_t *receivers, size_t n) {
    if (!receivers && n) return fail(p, "no receiver list");
    if (n > p->max_batch) return fail(p, "%zu receivers in one call, the pool was created for %zu", n, p->max_batch);
    for (size_t i = 0; i < n; i++) {
        if (receivers[i] >= p->rx.size()) return fail(p, "receiver %u out of range (%zu receivers)", receivers[i], p->rx.size());
        for (size_t k = 0; k < i; k++)
            if (receivers[k] == receivers[i]) return fail(p, "receiver %u listed twice: one buffer per receiver and call", receivers[i]);
    }
    return 0;
}

struct SinkAdapter { modes_pool_sink_fn fn; void *user; uint32_t receiver; };
void adapt(void *user, const modes_message *mm) {
    const SinkAdapter *a = static_cast<const SinkAdapter *>(user);
    if (a->fn) a->fn(a->user, a->receiver, mm);
}

int ensure_device(modes_pool *p) {
    if (p->ctx) return 0;
    modes_config c = p->cfg;
    c.n_gpus = 1; c.gpu_resolve = 0;
    p->ctx = modes_create(&c);
    if (!p->ctx) return fail(p, "%s", modes_last_error(nullptr));
    const size_t bytes = 2 * p->max_batch * kBuf;
    p->d_batch = static_cast<uint8_t *>(modes_device_alloc(bytes));
    p->h_carry = static_cast<uint8_t *>(modes_host_alloc(p->max_batch * MODES_CARRY_BYTES));
    if (!p->d_batch || !p->h_carry) return fail(p, "out of memory for %zu receivers per batch", p->max_batch);
    if (modes_device_memset(p->d_batch, 127, bytes)) return fail(p, "device memset failed");      // dump1090.c:344 no signal
    return 0;
}

}  // namespace

extern "C" {

modes_pool *modes_pool_create(const modes_config *cfg, size_t n_receivers, size_t max_batch_receivers) {
    if (!cfg || !n_receivers) return nullptr;
    modes_pool *p = new modes_pool();
    p->cfg = *cfg;
    p->max_batch = max_batch_receivers && max_batch_receivers < n_receivers ? max_batch_receivers : n_receivers;
    p->rx.resize(n_receivers);
    for (Receiver &r : p->rx) {
        r.res = modes_resolver_create(cfg);
        memset(r.carry, 127, sizeof r.carry);
        if (!r.res) { modes_pool_destroy(p); return nullptr; }
    }
    p->local_tiles.resize(modes_tile_count(1));
    return p;
}

void modes_pool_destroy(modes_pool *p) {
    if (!p) return;
    for (Receiver &r : p->rx) if (r.res) modes_resolver_destroy(r.res);
    if (p->d_batch) modes_device_free(p->d_batch);
    if (p->h_carry) modes_host_free(p->h_carry);
    if (p->ctx) modes_destroy(p->ctx);
    delete p;
}

const char *modes_pool_last_error(const modes_pool *p) { return p ? p->err.c_str() : "no pool"; }

int modes_pool_resolve(modes_pool *p, const uint32_t *receivers, size_t n, const modes_candidate *candidates,
                       const modes_tile *tiles, modes_pool_sink_fn sink, void *user) {
    if (!p) return -1;
    if (check_ids(p, receivers, n)) return -1;
    if (n && !tiles) return fail(p, "no tile table");
    const size_t n_tiles = modes_tile_count(2 * n);
    const size_t n_local = p->local_tiles.size();
    for (size_t i = 0; i < n; i++) {
        Receiver &r = p->rx[receivers[i]];
        // the data buffer of pair i is buffer 2i+1 of the batch; a position t belongs to the tile
        // that holds virtual position t + 2
        const int64_t t0 = kBufSamples * (int64_t)(2 * i + 1), t1 = t0 + kBufSamples;
        size_t g0 = (size_t)((t0 + 2) / MODES_TILE_SAMPLES), g1 = (size_t)((t1 + 1) / MODES_TILE_SAMPLES);
        if (g1 >= n_tiles) g1 = n_tiles - 1;
        p->local.clear();
        for (modes_tile &lt : p->local_tiles) { lt.offset = 0; lt.count = 0; }
        for (size_t g = g0; g <= g1; g++) {
            const modes_candidate *c = candidates + tiles[g].offset;
            for (uint32_t k = 0; k < tiles[g].count; k++) {
                if (c[k].t < t0 || c[k].t >= t1) continue;
                modes_candidate lc = c[k];
                lc.t -= t0;                               // position inside the receiver's own buffer
                const size_t lg = (size_t)((lc.t + 2) / MODES_TILE_SAMPLES);
                if (lg >= n_local) return fail(p, "candidate position %lld outside a buffer", (long long)lc.t);
                p->local_tiles[lg].count++;
                p->local.push_back(lc);
            }
        }
        uint32_t off = 0;
        for (modes_tile &lt : p->local_tiles) { lt.offset = off; off += lt.count; }
        SinkAdapter a{sink, user, receivers[i]};
        const size_t room = p->out && p->out_count < p->out_cap ? p->out_cap - p->out_count : 0;
        if (p->out) modes_resolver_set_output(r.res, room ? p->out + p->out_count : nullptr, room);
        if (modes_resolver_run(r.res, p->local.data(), p->local_tiles.data(), n_local, r.buffers, sink ? adapt : nullptr, &a))
            return fail(p, "resolve of receiver %u failed", receivers[i]);
        if (p->out) {
            const size_t k = modes_resolver_output_count(r.res);      // keeps counting past the room
            for (size_t m = 0; m < k && m < room; m++) p->out_rx[p->out_count + m] = receivers[i];
            p->out_count += k;
            modes_resolver_set_output(r.res, nullptr, 0);
        }
        r.buffers++;
    }
    return 0;
}

int modes_pool_ingest(modes_pool *p, const uint32_t *receivers, const uint8_t *const *iq, size_t n,
                      modes_pool_sink_fn sink, void *user) {
    if (!p) return -1;
    if (check_ids(p, receivers, n)) return -1;
    if (!n) return 0;
    if (!iq) return fail(p, "no buffers");
    if (ensure_device(p)) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(modes_stream(p->ctx));
    // the carries into the tails of the pad buffers, the new buffers behind them
    for (size_t i = 0; i < n; i++) memcpy(p->h_carry + i * MODES_CARRY_BYTES, p->rx[receivers[i]].carry, MODES_CARRY_BYTES);
    if (cudaMemcpy2DAsync(p->d_batch + kBuf - MODES_CARRY_BYTES, 2 * kBuf, p->h_carry, MODES_CARRY_BYTES, MODES_CARRY_BYTES, n,
                          cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail(p, "carry upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (size_t i = 0; i < n; i++) {
        if (!iq[i]) return fail(p, "receiver %u: no buffer", receivers[i]);
        if (cudaMemcpyAsync(p->d_batch + (2 * i + 1) * kBuf, iq[i], kBuf, cudaMemcpyHostToDevice, st) != cudaSuccess)
            return fail(p, "buffer upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (modes_detect_device(p->ctx, p->d_batch, 2 * n, nullptr, nullptr, 0, nullptr)) return fail(p, "%s", modes_last_error(p->ctx));
    uint64_t n_cand = 0;
    if (modes_detect_wait(p->ctx, &n_cand)) return fail(p, "%s", modes_last_error(p->ctx));
    p->cands.resize(n_cand ? n_cand : 1);
    p->tiles.resize(modes_tile_count(2 * n));
    if (modes_detect_fetch(p->ctx, p->cands.data(), p->tiles.data())) return fail(p, "%s", modes_last_error(p->ctx));
    // what each receiver carries into its next buffer (dump1090.c:481): taken before the resolve,
    // whose sink may hand the caller's buffers back
    for (size_t i = 0; i < n; i++)
        memcpy(p->rx[receivers[i]].carry, iq[i] + kBuf - MODES_CARRY_BYTES, MODES_CARRY_BYTES);
    return modes_pool_resolve(p, receivers, n, p->cands.data(), p->tiles.data(), sink, user);
}

int modes_pool_stats(const modes_pool *p, uint32_t receiver, modes_stats *out) {
    if (!p || receiver >= p->rx.size() || !out) return -1;
    return modes_resolver_stats(p->rx[receiver].res, out);
}

int modes_pool_reset(modes_pool *p, uint32_t receiver) {
    if (!p || receiver >= p->rx.size()) return -1;
    Receiver &r = p->rx[receiver];
    r.buffers = 0;
    memset(r.carry, 127, sizeof r.carry);
    return modes_resolver_reset(r.res);
}

int modes_pool_set_output(modes_pool *p, modes_message *out, uint32_t *receiver_of, size_t capacity) {
    if (!p || ((out == nullptr) != (receiver_of == nullptr))) return -1;
    p->out = out; p->out_rx = receiver_of; p->out_cap = out ? capacity : 0; p->out_count = 0;
    return 0;
}

size_t modes_pool_output_count(const modes_pool *p) { return p ? p->out_count : 0; }

int64_t modes_pool_buffers(const modes_pool *p, uint32_t receiver) {
    if (!p || receiver >= p->rx.size()) return -1;
    return p->rx[receiver].buffers;
}

}  // extern "C"
