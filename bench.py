#!/usr/bin/env python
"""bench.py — Msamples/s of the Mode S decode hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, all host cores

Workload (BASELINE.json configs[1]): testfiles/modes1.bin tiled back to back to 1 GiB
(536 870 912 samples = 4096 reference buffers) per GPU, --no-fix.  Weak scaling: rank r holds the
r-th GiB of the tiled stream.  A step = one pass of the hot path over the rank's GiB:
  value : inputs resident in HBM; scan + frame-evaluation kernels (+ NCCL gather of the candidate
          records to rank 0 when N>1); CUDA events on the launching stream, max over ranks.
  e2e   : the same GiB from pinned HOST memory through the public API to decoded messages on the
          host (H2D, kernels, D2H of records, sequential resolve on rank 0); wall clock between
          barriers + device synchronisation, max over ranks.
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GIB = 1 << 30
SAMPLES_PER_GIB = GIB // 2
METRIC = "Msamples/s (2 MHz u8 IQ) decoded, whole job"


def load_capture() -> tuple[np.ndarray, str]:
    """The reference's sample capture if it travelled with the repo, else a synthetic stand-in."""
    for p in (ROOT / "oracle" / "_ref" / "modes1.bin", Path("/root/reference/testfiles/modes1.bin")):
        if p.exists():
            return np.fromfile(p, dtype=np.uint8), "modes1.bin tiled"
    from dump1090_b200 import synth
    return synth.random_traffic(356868, 560, seed=1), "synthetic traffic (modes1.bin absent) tiled"


def shard_bytes(capture: np.ndarray, rank: int, nbytes: int = GIB) -> np.ndarray:
    """Bytes [rank*nbytes, (rank+1)*nbytes) of the capture tiled back to back."""
    start = (rank * nbytes) % capture.size
    reps = -(-(nbytes + start) // capture.size)
    return np.tile(capture, reps)[start: start + nbytes]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm

def _ref_worker(args):
    lo, hi, loops = args
    import checker
    return checker.ref_time(_REF_DATA[lo:hi], fix=0, loops=loops) if checker.REF_SO.exists() \
        else checker.oracle_time(_REF_DATA[lo:hi], fix=0, loops=loops)


_REF_DATA = None


def run_reference(args) -> None:
    """The reference's own computeMagnitudeVector + detectModeS loop (oracle/_ref, compiled from the
    unmodified sources) on the host cores: P processes on disjoint runs of whole buffers."""
    global _REF_DATA
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import checker
    kind = "reference" if checker.REF_SO.exists() else "port"
    if kind == "port":
        checker.build_oracle()
    capture, what = load_capture()
    cores = len(os.sched_getaffinity(0))
    sample_bytes = GIB
    _REF_DATA = shard_bytes(capture, 0, sample_bytes)
    nbuf = sample_bytes // 262144
    per = -(-nbuf // cores)
    slices = [(w * per * 262144, min(nbuf, (w + 1) * per) * 262144, 1) for w in range(cores) if w * per < nbuf]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(slices)) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, slices)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, slices)
        dt = time.perf_counter() - t0
    value = args.steps * (sample_bytes // 2) / dt / 1e6
    sample = f"{what} to 1 GiB per step (--no-fix), {len(slices)} processes on disjoint buffer runs"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic: " + what + " to 1 GiB per GPU",
        # same workload definition as the GPU arm; what one timed step of THIS arm covers is in `sample`
        "config": {"workload": "modes1.bin tiled to 1 GiB per GPU, --no-fix (BASELINE.json configs[1])",
                   "flags": "--no-fix", "samples_per_step": sample_bytes // 2},
        "cpu_baseline": {"value": round(value, 2), "unit": "Msamples/s", "cores": len(slices), "kind": kind,
                         "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- this framework

def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    from dump1090_b200 import api, sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # keep stdout clean for the one JSON line (NCCL prints its version banner there)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    numa = {"numa_node": None}
    if os.environ.get("BENCH_NUMA_BIND", "1") != "0":
        numa = sharded.bind_near_gpu(local_rank)              # before any pinned allocation
    if os.environ.get("BENCH_E2E_TRACE"):
        print(f"rank {rank}: numa binding {numa}", file=sys.stderr)

    capture, what = load_capture()
    nbuf = GIB // api.BUFFER_BYTES
    plan = [(r * nbuf, nbuf) for r in range(world)]
    pinned = api.PinnedBuffer(GIB)
    pinned.array[:] = shard_bytes(capture, rank)
    carry = None
    if rank > 0:
        prev = shard_bytes(capture, rank - 1)
        carry = bytes(prev[-api.CARRY_BYTES:])
    host_t = torch.from_numpy(pinned.array)
    d_iq = torch.empty(GIB, dtype=torch.uint8, device=dev)
    d_iq.copy_(host_t)
    cap = SAMPLES_PER_GIB // 64 + 4096
    d_cands = torch.empty(cap * 56, dtype=torch.uint8, device=dev)
    d_tiles = torch.empty(api.tiles_for(nbuf) * 8, dtype=torch.uint8, device=dev)

    dec = api.Decoder(fix_errors=0, device=local_rank, profile=1)
    stream = torch.cuda.Stream(device=dev)
    dec.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # N>1: the record gather is fused into the kernels — every rank's scan / frame-evaluation
    # kernels store their tile table and records straight into rank 0's HBM (CUDA IPC mapping,
    # NVLink stores); the only collective is a 4-byte all-reduce that orders "kernels done".
    pg = sharded.PeerGather(dist, rank, world, nbuf, cap) if world > 1 else None
    step_no = [0]
    fences = [None, None]

    def device_step():
        if world == 1:
            dec.detect_device(d_iq.data_ptr(), nbuf, carry, d_cands.data_ptr(), cap, d_tiles.data_ptr())
            return
        k = step_no[0] & 1
        step_no[0] += 1
        if fences[k] is not None:
            fences[k].wait()                           # stream-side: buffer k's previous round is complete
        pg.detect(dec, d_iq.data_ptr(), nbuf, carry, k)
        fences[k] = pg.fence()

    def device_drain():
        for f in fences:
            if f is not None:
                f.wait()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: inputs resident in HBM, device-timed
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            device_step()
        device_drain()
        barrier()
        dec.kernel_times_ms()                       # drop warm-up samples
        l0 = dec.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            device_step()
        device_drain()
        ev1.record(stream)
        barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = dec.launch_count() - l0
    ktimes = dec.kernel_times_ms()
    n_cand = dec.detect_wait()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * SAMPLES_PER_GIB * args.steps / (dev_ms * 1e-3) / 1e6

    # ---- e2e: host buffers in, messages out, through the public API
    e2e_msgs = 0
    d2h = 0
    if world == 1:
        dec2 = api.Decoder(fix_errors=0, device=local_rank)
        dec2.set_output_array(700000)

        def e2e_step():
            dec2.reset()
            dec2.rearm_output()
            dec2.process_ptr(pinned.ptr, GIB)
            dec2.finish()
            return dec2.output_count()
    else:
        # Steps are pipelined on rank 0: while its host threads resolve step i (records already in
        # its HBM, buffer i&1), every rank uploads and scans step i+1 into the other buffer.
        import threading
        resolver = None
        if rank == 0:
            resolver = api.Resolver(fix_errors=0)
            resolver.set_output_array(700000 * world)
        e2e_no = [0]
        worker = [None]
        result = [0]

        trace = os.environ.get("BENCH_E2E_TRACE") and rank == 0
        tr = []
        h2d_ms = []

        def resolve_async(k):
            t0 = time.perf_counter()
            shards = [(c, t, plan[r][0]) for r, (c, t) in enumerate(pg.fetch(k))]
            t1 = time.perf_counter()
            resolver.rearm_output()
            resolver.reset_state()
            resolver.run_shards(shards)
            result[0] = resolver.output_count()
            if trace:
                tr.append(("worker", round(1e3 * (t1 - t0), 2), round(1e3 * (time.perf_counter() - t1), 2)))

        # one resolver thread for the whole run, fed step numbers through a queue
        import queue
        jobs, done = queue.Queue(), queue.Queue()

        def resolver_loop():
            while True:
                k = jobs.get()
                if k is None:
                    return
                try:
                    resolve_async(k)
                    done.put(None)
                except BaseException as e:                   # surface it in e2e_join()
                    done.put(e)

        if rank == 0:
            threading.Thread(target=resolver_loop, daemon=True).start()

        def e2e_join():
            if worker[0] is not None:
                worker[0] = None
                err = done.get()
                if err is not None:
                    raise err
            return result[0]

        def e2e_step():
            k = e2e_no[0] & 1
            e2e_no[0] += 1
            t0 = time.perf_counter()
            pg.detect_host(dec, pinned.ptr, nbuf, carry, k)
            dec.detect_wait()
            t1 = time.perf_counter()
            h2d_ms.append(1e3 * (t1 - t0))
            pg.fence().wait()
            torch.cuda.synchronize(dev)
            t2 = time.perf_counter()
            if rank == 0:
                e2e_join()                                  # step i-1 resolved (its buffer is k^1)
                t3 = time.perf_counter()
                worker[0] = k
                jobs.put(k)
            else:
                t3 = t2
            if world > 1:
                dist.barrier()                              # nobody overwrites buffer k^1... see note
            if trace:
                tr.append(("step", round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2), round(1e3 * (t3 - t2), 2),
                           round(1e3 * (time.perf_counter() - t3), 2)))
            return result[0]

    with torch.cuda.stream(stream):
        for _ in range(min(args.warmup, 3)):
            e2e_msgs = e2e_step()
        if world > 1 and rank == 0:
            e2e_join()
        barrier()
        e2e_steps = args.steps
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_msgs = e2e_step()
        if world > 1 and rank == 0:
            e2e_msgs = e2e_join()
        barrier()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * SAMPLES_PER_GIB * e2e_steps / e2e_s / 1e6
    if world > 1 and os.environ.get("BENCH_E2E_TRACE"):
        print(f"rank {rank}: upload+kernels per step, ms: min {min(h2d_ms):.1f} median {sorted(h2d_ms)[len(h2d_ms) // 2]:.1f}", file=sys.stderr)
    if world > 1 and rank == 0 and os.environ.get("BENCH_E2E_TRACE"):
        print("e2e trace (ms): step = (h2d+kernels, fence+sync, join, barrier), worker = (fetch, resolve)", file=sys.stderr)
        for row in tr[-24:]:
            print("  ", row, file=sys.stderr)
    d2h = n_cand * 56 + api.tiles_for(nbuf) * 8 + 16

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        scan_ms = ktimes[0]
        achieved = 2.0 * SAMPLES_PER_GIB / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
        traffic = None
        tp = ROOT / "profiles" / "scan_kernel_traffic.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")

        cpu_baseline = None
        if world == 1:
            import checker
            kind = "reference" if checker.REF_SO.exists() else "port"
            if kind == "port":
                checker.build_oracle()
            sample = pinned.array[: 512 << 20]
            fn = checker.ref_time if kind == "reference" else checker.oracle_time
            secs = fn(sample, fix=0, loops=3)
            cpu_baseline = {"value": round(3 * (sample.size // 2) / secs / 1e6, 2), "unit": "Msamples/s", "cores": 1,
                            "kind": kind,
                            "sample": "first 512 MiB of the workload x 3 loops, single thread (the reference's "
                                      "own design: one decode thread), --no-fix"}
            if kind == "reference":
                try:                                    # the two hot calls separately; never fatal
                    part = sample[: 256 << 20]
                    t_mag, t_det = checker.ref_time_phases(part, fix=0, loops=1)
                    n = part.size // 2
                    cpu_baseline["phases"] = {"computeMagnitudeVector_Msamples_s": round(n / t_mag / 1e6, 1),
                                              "detectModeS_Msamples_s": round(n / t_det / 1e6, 1),
                                              "sample": "first 256 MiB, one pass"}
                except Exception as e:
                    cpu_baseline["phases"] = {"error": repr(e)}

        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic: " + what + " to 1 GiB per GPU",
            "config": {"workload": "modes1.bin tiled to 1 GiB per GPU, --no-fix (BASELINE.json configs[1])",
                       "flags": "--no-fix", "samples_per_step": world * SAMPLES_PER_GIB,
                       "candidates_per_gpu_step": n_cand,
                       "l2_policy": "input (1 GiB per GPU) is larger than the 126 MB L2; no explicit flush",
                       "step": "scan kernel (magnitude+preamble) + frame-evaluation kernel"
                               + (" with the record gather fused in: kernels store records into rank 0's HBM"
                                  " over NVLink (CUDA IPC), + a 4-byte NCCL all-reduce as completion fence"
                                  if world > 1 else "")},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": round(e2e_value, 1), "unit": "Msamples/s", "h2d_bytes_per_step": GIB + 480,
                    "d2h_bytes_per_step": int(d2h), "messages_per_step": int(e2e_msgs),
                    "ms_per_step": round(1e3 * e2e_s / e2e_steps, 3),
                    "path": "modes_process()+modes_finish() from pinned host memory" if world == 1 else
                            "H2D + modes_detect_device (records stored into rank 0's HBM) + fence + D2H + "
                            "modes_resolver_run_shards on rank 0"},
            "roofline": {"bound": "hbm", "kernel": "scan_kernel (fused magnitude + preamble tests)",
                         "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": 2 * SAMPLES_PER_GIB,
                         "scan_ms": round(scan_ms, 4), "eval_ms": round(ktimes[1], 4),
                         "batches_timed": int(ktimes[3])},
            "cpu_baseline": cpu_baseline,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        pg.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
