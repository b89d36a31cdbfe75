"""Multi-GPU decode: contiguous runs of whole reference buffers per rank.

The sample path shards with no data-path collective: a
candidate's evaluation depends only on samples [s-1, s+239] and on its index
inside its 131072-sample reference buffer, so rank r scans buffers
[r*B, (r+1)*B) given the 476 stream bytes that precede them.  The only
exchange is the gather of candidate records (56 bytes each) and tile tables to
rank 0, which replays the order-dependent half (retry/skip state machine, ICAO
cache — dump1090.c:1769-1791, :898-983) over the shards in stream order.

One process per GPU; torch.distributed supplies the gather (NCCL on GPUs; the
same code runs over gloo with CPU tensors for the host-logic tests).
"""
from __future__ import annotations

import numpy as np

from . import api


def shard_plan(total_buffers: int, world: int):
    """Contiguous, near-equal runs of whole buffers: [(first, count)] per rank."""
    base, extra = divmod(total_buffers, world)
    plan, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        plan.append((first, n))
        first += n
    return plan


def carry_before(stream: np.ndarray, first_buffer: int):
    """The 476 stream bytes preceding buffer `first_buffer` (None at stream start)."""
    if first_buffer == 0:
        return None
    off = first_buffer * api.BUFFER_BYTES
    return bytes(stream[off - api.CARRY_BYTES: off])


def gather_records(cands, tiles, n_cand: int, dist=None, group=None, dst: int = 0):
    """Gather per-rank (candidate records, tile table) to rank `dst`.

    cands: uint8 tensor holding >= n_cand*56 bytes (device tensor under NCCL, CPU
    tensor under gloo); tiles: uint8 tensor of the rank's tile table.  Returns on
    dst a list of (cands_bytes_tensor[:n*56], tiles_tensor) per rank, else None.
    Sizes are exchanged first so only max(n) records per rank travel.
    """
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [(cands[: n_cand * 56], tiles)]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    meta = torch.tensor([n_cand, tiles.numel()], dtype=torch.int64, device=cands.device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = [m.tolist() for m in metas]
    max_c = max(m[0] for m in metas) * 56
    max_t = max(m[1] for m in metas)
    send_c = cands[:max_c] if cands.numel() >= max_c else torch.cat(
        [cands, torch.zeros(max_c - cands.numel(), dtype=torch.uint8, device=cands.device)])
    send_t = tiles[:max_t] if tiles.numel() >= max_t else torch.cat(
        [tiles, torch.zeros(max_t - tiles.numel(), dtype=torch.uint8, device=tiles.device)])
    if rank == dst:
        rc = [torch.empty(max_c, dtype=torch.uint8, device=cands.device) for _ in range(world)]
        rt = [torch.empty(max_t, dtype=torch.uint8, device=tiles.device) for _ in range(world)]
        dist.gather(send_c.contiguous(), rc, dst=dst, group=group)
        dist.gather(send_t.contiguous(), rt, dst=dst, group=group)
        return [(rc[r][: metas[r][0] * 56], rt[r][: metas[r][1]]) for r in range(world)]
    dist.gather(send_c.contiguous(), None, dst=dst, group=group)
    dist.gather(send_t.contiguous(), None, dst=dst, group=group)
    return None


_pinned = {}


def _to_host(t, key):
    """Device tensor -> numpy view of a reusable pinned host buffer (async copy + one sync per call site)."""
    import torch
    if not t.is_cuda:
        return t.numpy()
    buf = _pinned.get(key)
    if buf is None or buf.numel() < t.numel():
        buf = torch.empty(max(t.numel(), 1), dtype=torch.uint8, pin_memory=True)
        _pinned[key] = buf
    buf[: t.numel()].copy_(t, non_blocking=True)
    return buf[: t.numel()].numpy()


def resolve_gathered(resolver, gathered, plan) -> None:
    """Rank 0: replay the sequential half over all shards (stream order).  The shards are resolved
    concurrently on host threads; modes_resolver_run_shards keeps that exact by verifying the
    speculated ICAO-cache state at every shard boundary."""
    import torch
    host = [(_to_host(c, ("c", r)), _to_host(t, ("t", r))) for r, (c, t) in enumerate(gathered)]
    if any(c.is_cuda for c, _ in gathered):
        torch.cuda.current_stream().synchronize()
    shards = []
    for (c, t), (first, _n) in zip(host, plan):
        c_np = c.view(api.CANDIDATE_DTYPE) if c.size else np.zeros(0, dtype=api.CANDIDATE_DTYPE)
        shards.append((c_np, t.view(api.TILE_DTYPE), first))
    resolver.run_shards(shards)


class ShmExchange:
    """all_gather of small uint32 vectors between the ranks of ONE node through a file in /dev/shm:
    every rank owns two rows [seq, size, payload...], writes its payload into the row of the
    exchange's parity, then its sequence number, and polls the others' — microseconds instead of the
    milliseconds of a gloo ring over loopback (measured: the resolve worker of a 4-rank job spent a
    third of its time in three gloo all-gathers of 8 KiB).  Two rows make one wait per exchange
    enough: a rank overwrites the row of exchange s only in exchange s+2, which it enters after it
    has seen every rank publish s+1, and a rank publishes s+1 only after it has read everybody's s.
    Create it collectively (the name travels by broadcast_object_list)."""

    def __init__(self, dist, group=None, max_words: int = 2 * 1024):
        import os
        import tempfile
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.words = max_words
        name = [None]
        if self.rank == 0:
            fd, path = tempfile.mkstemp(prefix="modes_b200_xchg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            os.ftruncate(fd, self.world * 2 * (2 + max_words) * 4)
            os.close(fd)
            name[0] = path
        dist.broadcast_object_list(name, src=0, group=group)
        self.path = name[0]
        self.mem = np.memmap(self.path, dtype=np.uint32, mode="r+", shape=(self.world, 2, 2 + max_words))
        self.seq = 0
        dist.barrier(group=group)

    def all_gather(self, vec: np.ndarray) -> np.ndarray:
        import time
        assert vec.size <= self.words
        self.seq += 1
        bank = self.seq & 1
        row = self.mem[self.rank, bank]
        row[2: 2 + vec.size] = vec
        row[1] = vec.size
        row[0] = self.seq                                   # published last
        out = np.empty((self.world, vec.size), dtype=np.uint32)
        deadline = time.perf_counter() + 120.0
        for r in range(self.world):
            self._await(r, bank, deadline)
            out[r] = self.mem[r, bank, 2: 2 + vec.size]
        return out

    def _await(self, r: int, bank: int, deadline: float):
        """Poll rank r's sequence number: yield for the first 200 us (ranks in step arrive within
        microseconds), then sleep between polls — a rank that is a whole upload late must not cost
        the waiting ranks a core each."""
        import time
        t0 = time.perf_counter()
        while self.mem[r, bank, 0] < self.seq:
            now = time.perf_counter()
            if now > deadline:
                raise TimeoutError(f"ShmExchange: rank {r} did not arrive at exchange {self.seq}")
            time.sleep(0 if now - t0 < 200e-6 else 100e-6)

    def close(self):
        import os
        self.mem = None
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def resolve_distributed(resolver, cands: np.ndarray, tiles: np.ndarray, buffer_base: int, dist=None, group=None,
                        start_cache=None, exchange=None) -> dict:
    """Every rank resolves ITS OWN shard (records in its own host memory) and the result is exactly
    the sequential one: no rank sees another rank's records, only 4 KiB address caches travel.

    Protocol (include/modes_b200.h, "one shard of a sharded decode"): all-gather what the tail of
    each shard leaves in an empty cache; rank k guesses its starting cache as the job's starting
    cache overwritten by shard k-1's tail; every rank resolves tentatively; all-gather (guess, end
    cache); a rank is verified when its guess equals what the previous rank really ended with and
    that rank is verified; unverified ranks run again from the now known cache.  One round is the
    rule, world rounds the worst case.  Then every rank commits (delivers) its own messages.
    `group` must accept CPU tensors (gloo); `exchange` (ShmExchange) replaces it for the 4 KiB
    vectors on one node.  Returns {"rounds", "end_cache" (of the whole job), "ms": this rank's time in
    the tail pass, the exchanges (waiting for the slowest rank included), the tentative runs, the commit}."""
    import torch
    world = 1 if dist is None or not dist.is_initialized() else dist.get_world_size(group)
    rank = 0 if world == 1 else dist.get_rank(group)
    S0 = np.zeros(api.ICAO_CACHE_SLOTS, dtype=np.uint32) if start_cache is None else np.asarray(start_cache, dtype=np.uint32)
    cands = np.ascontiguousarray(cands)
    tiles = np.ascontiguousarray(tiles)

    def all_gather(vec: np.ndarray) -> np.ndarray:
        if world == 1:
            return vec[None, :]
        if exchange is not None:
            return exchange.all_gather(vec)
        t = torch.from_numpy(vec.astype(np.int64))            # gloo has no uint32
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        return np.stack([o.numpy().astype(np.uint32) for o in outs])

    import time
    clock = time.perf_counter
    ms = {"tail": 0.0, "exchange": 0.0, "tentative": 0.0, "commit": 0.0}    # where this call's time went

    def timed(key, fn, *a):
        t0 = clock()
        r = fn(*a)
        ms[key] += 1e3 * (clock() - t0)
        return r

    tails = None
    if world > 1:
        mine = timed("tail", resolver.tail_cache, cands, tiles, buffer_base)
        tails = timed("exchange", all_gather, mine)          # includes waiting for the slowest rank to get here
    guess = S0.copy()
    if rank > 0:
        t = tails[rank - 1]
        guess[t != 0] = t[t != 0]
    rounds, need_run, end = 0, True, None
    while True:
        if need_run:
            resolver.set_cache(guess)
            timed("tentative", resolver.run_tentative, cands, tiles, buffer_base)
            end = resolver.get_cache()
        rounds += 1
        both = timed("exchange", all_gather, np.concatenate([guess, end]))
        G, E = both[:, : api.ICAO_CACHE_SLOTS], both[:, api.ICAO_CACHE_SLOTS:]
        ok = [bool(np.array_equal(G[0], S0))]
        for k in range(1, world):
            ok.append(ok[k - 1] and bool(np.array_equal(G[k], E[k - 1])))
        if all(ok):
            break
        # everything up to the first unverified rank is final; from there on, run again from the
        # cache the predecessor ended with in this round (right for the first of them at least)
        need_run = not ok[rank]
        if need_run:
            guess = E[rank - 1].copy() if rank > 0 else S0.copy()
    timed("commit", resolver.commit)
    return {"rounds": rounds, "end_cache": E[world - 1].copy(), "ms": ms}


def gather_fixed(cands, tiles, dist, group=None, dst: int = 0, out=None):
    """Gather fixed-size record/tile buffers to rank `dst` with no host round trip (counts travel
    in the tile tables).  `out` = preallocated ([world x cands], [world x tiles]) on dst.  Returns
    the two async work handles."""
    rank = dist.get_rank(group)
    w1 = dist.gather(cands, out[0] if rank == dst else None, dst=dst, group=group, async_op=True)
    w2 = dist.gather(tiles, out[1] if rank == dst else None, dst=dst, group=group, async_op=True)
    return w1, w2


def bind_near_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned
    host buffers it allocates next are local to that GPU's PCIe root (a rank whose buffers sit on
    the other socket uploads at a fraction of the link rate, and a sharded step is as slow as its
    slowest rank).  Best effort: returns what it did, never raises."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:                                    # no sysfs, no permission, ...: stay unbound
        info["error"] = repr(e)
    return info


class PeerGather:
    """Record gather fused into the kernels: every rank's scan / frame-evaluation kernels store
    their tile table and candidate records straight into a buffer in rank 0's HBM (mapped through
    CUDA IPC, written over NVLink), so no collective follows the compute.  The only collective left
    is a 4-byte all-reduce that tells rank 0 the peers' kernels have finished.

    Layout of one buffer: world segments of [16-byte header {found, overflow, 0, 0}][tile table]
    [records]; `nbuf` double-buffers so that rank 0 can read step i while step i+1 is written.
    """

    def __init__(self, dist, rank: int, world: int, n_buffers_per_rank: int, cap: int, nbuf: int = 2, group=None):
        import torch
        self.dist, self.rank, self.world, self.cap, self.group = dist, rank, world, cap, group
        self.n_tiles = api.tiles_for(n_buffers_per_rank)
        self.tiles_off = 16
        self.rec_off = (16 + self.n_tiles * 8 + 255) // 256 * 256
        self.seg = (self.rec_off + cap * 56 + 255) // 256 * 256
        self.buf_bytes = self.seg * world
        self.nbuf = nbuf
        L = api.lib()
        handle = torch.zeros(64, dtype=torch.uint8)
        self._own = None
        if rank == 0:
            self._own = L.modes_device_alloc(self.buf_bytes * nbuf)
            if not self._own:
                raise MemoryError("modes_device_alloc failed")
            hb = (api.C.c_uint8 * 64)()
            if L.modes_ipc_export(api.C.c_void_p(self._own), hb):
                raise RuntimeError("modes_ipc_export failed")
            handle = torch.tensor(list(hb), dtype=torch.uint8)
        dev = torch.device("cuda", torch.cuda.current_device())
        h = handle.to(dev)
        dist.broadcast(h, src=0, group=group)
        if rank == 0:
            self.base = self._own
        else:
            hb = (api.C.c_uint8 * 64)(*h.cpu().tolist())
            self.base = L.modes_ipc_open(hb)
            if not self.base:
                raise RuntimeError("modes_ipc_open failed (no peer access to rank 0?)")
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._pin = None
        if rank == 0:                                   # pinned landing zone for fetch()
            self._pin = api.PinnedBuffer(world * (self.n_tiles * 8 + cap * 56))

    def segment(self, k: int, rank: int | None = None):
        """(header_ptr, tiles_ptr, records_ptr) of `rank`'s segment in buffer k."""
        r = self.rank if rank is None else rank
        b = self.base + (k % self.nbuf) * self.buf_bytes + r * self.seg
        return b, b + self.tiles_off, b + self.rec_off

    def detect(self, dec, d_iq_ptr: int, n_buffers: int, carry, k: int) -> None:
        """Launch this rank's kernels with their outputs in rank 0's memory, then publish the count."""
        hdr, tiles, recs = self.segment(k)
        dec.detect_device(d_iq_ptr, n_buffers, carry, recs, self.cap, tiles)
        dec.publish_count(hdr)

    def detect_host(self, dec, host_ptr: int, n_buffers: int, carry, k: int) -> None:
        """Same from (pinned) host memory: H2D + kernels on the decoder's stream."""
        hdr, tiles, recs = self.segment(k)
        dec.detect_host(host_ptr, n_buffers, carry, recs, self.cap, tiles)
        dec.publish_count(hdr)

    def fence(self):
        """All ranks: after this (stream-ordered) every rank's kernels up to here have completed."""
        return self.dist.all_reduce(self._flag, group=self.group, async_op=True)

    def fetch(self, k: int):
        """Rank 0, after fence().wait() and a stream sync: [(cands ndarray, tiles ndarray)] per rank."""
        import torch
        assert self.rank == 0
        out = []
        hdrs = np.zeros((self.world, 4), dtype=np.uint32)
        L = api.lib()
        def d2h(dst, src, n):
            if L.modes_copy_to_host(api.C.c_void_p(dst), api.C.c_void_p(src), n):
                raise RuntimeError("device-to-host copy failed")
        for r in range(self.world):
            hdr, tiles, recs = self.segment(k, r)
            d2h(hdrs[r].ctypes.data, hdr, 16)
        for r in range(self.world):
            n, ovf = int(hdrs[r][0]), int(hdrs[r][1])
            if ovf or n > self.cap:
                raise RuntimeError(f"rank {r}: candidate capacity exceeded ({n} > {self.cap})")
            hdr, tiles, recs = self.segment(k, r)
            off = r * (self.n_tiles * 8 + self.cap * 56)
            t = self._pin.array[off: off + self.n_tiles * 8].view(api.TILE_DTYPE)
            c = self._pin.array[off + self.n_tiles * 8: off + self.n_tiles * 8 + max(n, 1) * 56].view(api.CANDIDATE_DTYPE)
            d2h(t.ctypes.data, tiles, t.nbytes)
            if n:
                d2h(c.ctypes.data, recs, n * 56)
            out.append((c[:n], t))
        return out

    def close(self):
        L = api.lib()
        if self.rank == 0 and self._own:
            L.modes_device_free(self._own)
            self._own = None
            self._pin.free()
        elif self.rank != 0 and self.base:
            L.modes_ipc_close(self.base)
            self.base = None
