#!/usr/bin/env python3
"""bench.py - the hot path's headline metric on B200.

Metric (BASELINE.json): ZIP/BGZF inflate GB/s of decompressed output.  Default
workload = config "zip64k": a ZIP of 4096 x 64 KiB synthetic-text entries,
deflate level 6 (256 MiB out per GPU), located through the central directory,
every entry inflated and CRC-32 verified in ONE device pass per step.

  value    device-resident: archive already in HBM, K timed passes of
           b2i_plan_launch (CUDA events on the launching stream)
  e2e      the same archive through the reference-facing C ABI call
           b2i_decode_host with PINNED HOST buffers: descriptor upload + H2D of
           the compressed span + kernel + D2H of the decoded bytes + results,
           all inside the timed region
  roofline inflate kernel: algorithmic bytes (csize + usize per stream) per
           launch / mean launch duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified
           libarchive + zlib) on the box's host cores, bounded sample

Multi-GPU (--gpus N, launched by torchrun): entries shard by rank, every rank
decodes its own archive of the same shape (weak scaling), no data-path
collective; torch.distributed (NCCL) is used only for the barrier and the
max-over-ranks of the device time.

`--impl reference` times the reference CPU implementation (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "zip64k": "ZIP of 4096 x 64 KiB synthetic-text entries, deflate level 6 (256 MiB out), CRC check on",
    "stored1m": "ZIP of 1024 x 1 MiB stored entries: CRC-32 verification only",
    "bgzf64k": "BGZF multi-member gzip, 64 KiB members (scaled: 16384 members = 1 GiB out)",
    "tiny4k": "ZIP64 of 4 KiB text entries (scaled: 65536 entries = 256 MiB)",
    "mixed": "ZIP64, log-uniform entry sizes 1 KiB-16 MiB, dynamic/fixed/stored blocks mixed (scaled: 1 GiB out)",
}


def build_workload(name, rank, scale=1.0):
    from libarchive_b200 import synth
    threads = max(4, min(32, (os.cpu_count() or 8)))
    if name == "zip64k":
        n = max(8, int(4096 * scale))
        parts = synth.split_text(n * 65536, 65536, 12345 + rank)
        return synth.make_zip([synth.ZipMember("e%06d.txt" % i, p) for i, p in enumerate(parts)],
                              threads=threads), "zip"
    if name == "stored1m":
        n = max(2, int(1024 * scale))
        blob = synth.synth_random(n << 20, 2 + rank)
        return synth.make_zip([synth.ZipMember("s%05d.bin" % i, blob[i << 20:(i + 1) << 20], method=0)
                               for i in range(n)], threads=1), "zip"
    if name == "bgzf64k":
        n = max(8, int(16384 * scale))
        parts = synth.split_text(n * 65280, 65280, 54321 + rank)
        return synth.make_bgzf(parts, threads=threads), "bgzf"
    if name == "tiny4k":
        n = max(64, int(65536 * scale))
        parts = synth.split_text(n * 4096, 4096, 5 + rank)
        return synth.make_zip([synth.ZipMember("t%06d" % i, p) for i, p in enumerate(parts)],
                              zip64=True, threads=threads), "zip"
    if name == "mixed":
        return synth.config4_zip64_mixed(total=max(8 << 20, int((1 << 30) * scale)), seed=4 + rank), "zip"
    raise SystemExit("unknown workload " + name)


def plan_for(archive, kind):
    from libarchive_b200 import capi, reader
    if kind == "zip":
        entries, _, _ = capi.zip_index(archive)
        descs, out_bytes, which = reader.plan_zip(entries)
    else:
        members, _ = capi.gzip_scan_bgzf(archive)
        descs, out_bytes = reader.plan_bgzf(members)
    usize = sum(int(d.expect_out) for d in descs)
    csize = sum(int(d.in_len) for d in descs)
    return descs, out_bytes, usize, csize


def algorithmic_bytes(descs):
    """HBM bytes one launch must move: csize read + usize written per
    inflated stream; n bytes read per stored entry verified in place (nothing written)."""
    from libarchive_b200 import capi
    total = 0
    for d in descs:
        total += int(d.in_len)
        if not (d.method == 0 and d.flags & capi.F_NO_COPY):
            total += int(d.expect_out)
    return total


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- CPU reference arm

def cpu_reference_run(archive, kind, steps, warmup, sample_units=None, procs=None):
    """Time the reference's own CPU implementation (oracle/_ref = unmodified libarchive +
    zlib; falls back to the oracle port when _ref is absent) on a bounded sample."""
    from libarchive_b200 import capi
    nproc = procs or (os.cpu_count() or 1)
    extract = os.path.join(ROOT, "oracle", "_ref", "oracle_extract")
    if kind == "zip":
        entries, _, _ = capi.zip_index(archive)
        total_units = len(entries)
    else:
        members, _ = capi.gzip_scan_bgzf(archive)
        total_units = len(members)
    units = min(total_units, sample_units or 4096)
    if os.path.exists(extract):
        with tempfile.NamedTemporaryFile(prefix="b2i_cpu_", suffix=".bin", delete=False, dir="/tmp") as f:
            f.write(archive)
            path = f.name
        try:
            cmd = [extract, "bench", path, "--procs", str(nproc), "--reps", "3", "--limit", str(units)]
            if kind != "zip":
                cmd.append("--raw")
            times, out_bytes = [], 0
            for i in range(warmup + steps):
                r = subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=600)
                j = json.loads(r.stdout.strip().splitlines()[-1])
                if j.get("bad"):
                    raise RuntimeError("reference reported errors: %s" % j)
                if i >= warmup:
                    times.append(j["seconds"])
                out_bytes = j["out_bytes"]
        finally:
            os.unlink(path)
        sec = sum(times) / len(times)
        return {"value": out_bytes / sec / 1e9, "unit": "GB/s", "cores": nproc, "kind": "reference",
                "sample": "%d of %d %s via oracle/_ref (unmodified libarchive + zlib %s), one process per core, "
                          "archive_read_data into 64 KiB buffer, CRC on" %
                          (units, total_units, "entries" if kind == "zip" else "members", j.get("zlib", "?")),
                "seconds": sec, "out_bytes": out_bytes}
    # oracle port (scalar, 1 thread)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding as ob
    descs, out_bytes, usize, csize = plan_for(archive, kind)
    units = min(len(descs), 64)
    sub = (type(descs[0]) * units)(*descs[:units])
    need = max(int(d.out_off + d.out_cap) for d in sub)
    t0 = time.perf_counter()
    res, _ = ob.decode_batch(archive, sub, need)
    sec = time.perf_counter() - t0
    nbytes = sum(int(r.out_bytes) for r in res)
    return {"value": nbytes / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "%d streams through the scalar oracle port" % units, "seconds": sec, "out_bytes": nbytes}


# --------------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="zip64k", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (tests only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the CRC-only config (stored1m)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = args.steps

    if args.impl == "reference":
        if rank != 0:
            return 0
        archive, kind = build_workload(args.workload, 0, args.scale)
        cb = cpu_reference_run(archive, kind, K, W)
        line = {"metric": "inflate_out_GBps", "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference",
                "config": {"workload": args.workload, "description": WORKLOADS[args.workload]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from libarchive_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    # one process per GPU: staging buffers on the GPU's own NUMA node (multi-socket boxes)
    from libarchive_b200.shard import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 and not os.environ.get("B2I_NO_NUMA_BIND") else 0
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    archive, kind = build_workload(args.workload, rank, args.scale)
    descs, out_bytes, usize, csize = plan_for(archive, kind)
    n = len(descs)
    # a dedicated stream shared by torch (events) and the library (kernels, copies);
    # the legacy default stream has handle 0, which the C ABI reads as "make your own"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local_rank, stream.cuda_stream)
    L = capi.lib()

    # pinned host buffers (the plugin's staging) and device-resident copies
    in_bytes = len(archive)
    h_in = L.b2i_host_alloc(in_bytes + 64)
    h_out = L.b2i_host_alloc(out_bytes + 64)
    h_out2 = L.b2i_host_alloc(out_bytes + 64)      # second job in flight (b2i_submit / b2i_wait)
    C.memmove(h_in, archive, in_bytes)
    d_in = L.b2i_device_alloc(ctx.h, in_bytes + 64)
    d_out = L.b2i_device_alloc(ctx.h, out_bytes + 64)
    assert h_in and h_out and h_out2 and d_in and d_out
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, h_in, in_bytes))
    plan = C.c_void_p()
    ctx._check(L.b2i_plan_create(ctx.h, descs, n, C.byref(plan)))
    res = (capi.StreamResult * n)()

    def device_step():
        ctx._check(L.b2i_plan_launch(plan, d_in, in_bytes, d_out, out_bytes))

    def e2e_step():
        ctx._check(L.b2i_decode_host(ctx.h, h_in, in_bytes, descs, n, h_out, out_bytes, res))

    # correctness gate before timing: every stream OK, sizes and CRCs as the directory says
    device_step()
    ctx._check(L.b2i_plan_results(plan, res))
    bad = [i for i in range(n) if res[i].status != 0 or res[i].flags != 0]
    if bad:
        raise SystemExit("bench.py: %d streams failed verification (first: %d status %d flags %d)" %
                         (len(bad), bad[0], res[bad[0]].status, res[bad[0]].flags))

    # ---- device-resident timing --------------------------------------------------
    for _ in range(W):
        device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    evs[0].record(stream)
    for i in range(K):
        device_step()
        evs[i + 1].record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = evs[0].elapsed_time(evs[K])

    # ---- host link: pinned H2D / D2H bandwidth for this workload's sizes ------------
    def link_gbs(fn, nbytes, reps=5):
        best = 0.0
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        return best
    h2d_gbs = link_gbs(lambda: ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, h_in, in_bytes)), in_bytes)
    d2h_gbs = link_gbs(lambda: ctx._check(L.b2i_memcpy_d2h(ctx.h, h_out, d_out, max(out_bytes, 1))), max(out_bytes, 1))
    # a serial copy-in + copy-out of this step's bytes is the host-link floor of one step
    link_floor_ms = (csize / (h2d_gbs * 1e9) + out_bytes / (d2h_gbs * 1e9)) * 1e3
    overlap_floor_ms = max(csize / (h2d_gbs * 1e9), out_bytes / (d2h_gbs * 1e9)) * 1e3

    # ---- end-to-end timing (host buffers, copies inside) -------------------------
    for _ in range(min(W, 3)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_sync_ms = max(e0.elapsed_time(e1), e2e_wall_ms)   # host API is synchronous: wall covers host work too
    bad = [i for i in range(n) if res[i].status != 0 or res[i].flags != 0]
    if bad:
        raise SystemExit("bench.py: e2e pass failed verification for %d streams" % len(bad))

    # The same K steps with two jobs in flight (b2i_submit / b2i_wait): every step still
    # copies its input from pinned host memory and its whole output back; the copy-out of
    # step i overlaps the copy-in and decode of step i+1.  Wall clock: a step is done when
    # b2i_wait has returned, i.e. its output and results are in host memory.
    def e2e_pipelined(steps):
        jobs, last = [], None
        for i in range(steps):
            if len(jobs) == 2:
                last = ctx.wait(jobs.pop(0))
            jobs.append(ctx.submit(h_in, in_bytes, descs, h_out if i % 2 == 0 else h_out2, out_bytes))
        outs = [last] if last is not None else []
        while jobs:
            outs.append(ctx.wait(jobs.pop(0)))
        return outs

    e2e_pipelined(min(W, 3))
    barrier()
    t0 = time.perf_counter()
    outs = e2e_pipelined(K)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop()
    for r_ in outs:
        bad = [i for i in range(n) if r_[i].status != 0 or r_[i].flags != 0]
        if bad:
            raise SystemExit("bench.py: pipelined e2e pass failed verification for %d streams" % len(bad))
    if C.string_at(h_out, min(out_bytes, 1 << 20)) != C.string_at(h_out2, min(out_bytes, 1 << 20)):
        raise SystemExit("bench.py: the two jobs in flight produced different output")

    t = torch.tensor([total_ms, e2e_ms, float(usize), float(csize), e2e_sync_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms, e2e_sync_ms = float(tmax[0]), float(tmax[1]), float(tmax[4])
        all_usize, all_csize = float(tsum[2]), float(tsum[3])
    else:
        all_usize, all_csize = float(usize), float(csize)

    ms_per_step = total_ms / K
    value = all_usize * K / (total_ms * 1e-3) / 1e9
    e2e_value = all_usize * K / (e2e_ms * 1e-3) / 1e9
    peak, peak_src = measured_hbm_peak()
    kern_ms = statistics.mean(step_ms)
    alg_bytes = algorithmic_bytes(descs)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")

    line = {
        "metric": "inflate_out_GBps" if args.workload != "stored1m" else "crc32_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "description": WORKLOADS[args.workload],
                   "streams_per_gpu": n, "out_bytes_per_gpu": usize, "in_bytes_per_gpu": csize,
                   "l2": "working set (in+out %.0f MB) exceeds the 126 MB L2; no flush" % ((usize + csize) / 1e6),
                   "parallelism": "entries sharded by rank, no collective", "scale": args.scale, "numa_bound_cpus": numa_cpus},
        "roofline": {"bound": "hbm", "kernel": "b2i_inflate_kernel" if args.workload != "stored1m" else "b2i_crc_chunks_kernel",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern_ms,
                     "out_frac_of_hbm": (usize / (kern_ms * 1e-3) / 1e9) / peak},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": csize + n * 48,
                "d2h_bytes_per_step": out_bytes + n * 32, "ms_per_step": e2e_ms / K,
                "mode": "b2i_submit/b2i_wait, two jobs in flight; every step copies its input from pinned host "
                        "memory and its whole output back",
                "one_call_at_a_time": {"value": all_usize * K / (e2e_sync_ms * 1e-3) / 1e9, "unit": "GB/s",
                                       "ms_per_step": e2e_sync_ms / K, "call": "b2i_decode_host"},
                "host_link": {"h2d_GBps": h2d_gbs, "d2h_GBps": d2h_gbs,
                              "frac_of_link": overlap_floor_ms / (e2e_ms / K),
                              "note": "frac = time the slower direction alone needs at the measured pinned "
                                      "bandwidth / measured e2e step time"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    if args.extra and rank == 0:
        line["crc32"] = crc_config(ctx, L, stream, torch, K, W, peak)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_run(archive, kind, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            one = cpu_reference_run(archive, kind, 1, 0, sample_units=512, procs=1)
            line["cpu_baseline"]["one_core_value"] = one["value"]
        except Exception as ex:  # the baseline must not sink the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
    elif rank == 0:
        line["cpu_baseline"] = None

    L.b2i_plan_destroy(plan)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def crc_config(ctx, L, stream, torch, K, W, peak):
    """BASELINE config 2: 1024 x 1 MiB stored entries, CRC-32 verification throughput."""
    from libarchive_b200 import capi
    archive, kind = build_workload("stored1m", 0)
    descs, out_bytes, usize, csize = plan_for(archive, kind)
    n = len(descs)
    d_in = L.b2i_device_alloc(ctx.h, len(archive) + 64)
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, archive, len(archive)))
    plan = C.c_void_p()
    ctx._check(L.b2i_plan_create(ctx.h, descs, n, C.byref(plan)))
    res = (capi.StreamResult * n)()
    for _ in range(W):
        ctx._check(L.b2i_plan_launch(plan, d_in, len(archive), None, 0))
    ctx._check(L.b2i_plan_results(plan, res))
    assert all(r.status == 0 and r.flags == 0 for r in res)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(K):
        ctx._check(L.b2i_plan_launch(plan, d_in, len(archive), None, 0))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    L.b2i_plan_destroy(plan)
    L.b2i_device_free(ctx.h, d_in)
    gbs = csize / (ms * 1e-3) / 1e9
    return {"workload": "stored1m", "value": gbs, "unit": "GB/s", "ms_per_step": ms, "frac_of_hbm": gbs / peak,
            "bytes": csize}


if __name__ == "__main__":
    sys.exit(main())
