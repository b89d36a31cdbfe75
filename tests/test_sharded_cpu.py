"""CPU (-m "not gpu"): the multi-GPU host logic over gloo, world size 2.

The device half is replaced by the checker's candidate records (oracle_scan_candidates); what is
under test is the product's sharding plan, the record gather and the stream-ordered resolve on
rank 0 (dump1090_b200/sharded.py + modes_resolver_*)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import checker as C
from dump1090_b200 import api, sharded, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nsamples, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = synth.random_traffic(nsamples, nsamples // 700, seed)
        total_buffers = data.size // api.BUFFER_BYTES + 1
        plan = sharded.shard_plan(total_buffers, world)
        first, count = plan[rank]
        # stand-in for the device half: this rank's candidates, shard-local positions
        cands = C.oracle_scan_candidates(data)
        arr = np.frombuffer(b"".join(bytes(c) for c in cands), dtype=api.CANDIDATE_DTYPE).copy()
        mine = arr[((arr["t"] >> 17) >= first) & ((arr["t"] >> 17) < first + count)].copy()
        mine["t"] -= first << 17
        # two tiles per shard to exercise the tile walk
        half = mine.size // 2
        tiles = np.array([(0, half), (half, mine.size - half)], dtype=api.TILE_DTYPE)
        c_t = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()) if mine.size else torch.zeros(0, dtype=torch.uint8)
        t_t = torch.from_numpy(tiles.view(np.uint8).reshape(-1).copy())
        gathered = sharded.gather_records(c_t, t_t, mine.size, dist)
        if rank == 0:
            res = api.Resolver()
            res.set_output_array(100000)
            sharded.resolve_gathered(res, gathered, plan)
            n = res.output_count()
            lines = [res._out[i].raw_line() for i in range(n)]
            exp, exp_stats = C.oracle_decode(data)
            q.put((lines == [m.hexline() for m in exp], list(res.stats().values()) == exp_stats, n))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nsamples,seed", [(600000, 31), (262144 * 2 + 1000, 32)])
def test_two_rank_gather_and_resolve(nsamples, seed, checker_libs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nsamples, seed, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    same_lines, same_stats, n = q.get(timeout=10)
    assert same_lines and same_stats and n > 0


def test_shard_plan_and_carry():
    assert sharded.shard_plan(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sharded.shard_plan(2, 4) == [(0, 1), (1, 1), (2, 0), (2, 0)]
    stream = (np.arange(3 * api.BUFFER_BYTES) % 251).astype(np.uint8)
    assert sharded.carry_before(stream, 0) is None
    c = sharded.carry_before(stream, 2)
    assert len(c) == api.CARRY_BYTES and c == bytes(stream[2 * api.BUFFER_BYTES - api.CARRY_BYTES: 2 * api.BUFFER_BYTES])


def test_bind_near_gpu_is_best_effort():
    """No GPU (or no sysfs entry) must not raise: the helper reports and leaves the affinity alone."""
    import os
    from dump1090_b200 import sharded
    before = os.sched_getaffinity(0)
    info = sharded.bind_near_gpu(0)
    assert "numa_node" in info and "cpus" in info
    import torch
    if not torch.cuda.is_available():
        assert info["cpus"] is None and os.sched_getaffinity(0) == before
