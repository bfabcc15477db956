#!/usr/bin/env python3
"""bench.py - the hot path's headline metric on B200, every BASELINE config in one line.

Metric (BASELINE.json): ZIP/BGZF inflate GB/s of decompressed output; CRC-32 GB/s.
Headline workload = config 1 "zip64k": a ZIP of 4096 x 64 KiB synthetic-text entries,
deflate level 6 (256 MiB out per GPU), located through the central directory, every
entry inflated and CRC-32 verified in ONE device pass per step.

  value      device-resident: archive already in HBM, K timed passes of b2i_plan_launch
             (CUDA events on the launching stream)
  e2e        the same archive through libarchive's PUBLIC read API on the drop-in
             library (archive_read_open_memory + archive_read_next_header +
             archive_read_data_block; a separate process, source bytes in pageable host
             memory, every byte crosses the host link inside the timed region); the
             archive_read_data(64 KiB buffer) figure and the C-ABI figures
             (b2i_decode_host, b2i_submit/b2i_wait, b2i_pipe) sit beside it
  roofline   inflate kernel: algorithmic bytes (csize + usize per stream) per launch /
             mean launch duration, against MEASURED_PEAKS.json hbm_gbs
  configs    configs 2-5 (stored1m / bgzf64k / mixed / tiny4k): device-resident GB/s,
             roofline fraction, C-ABI e2e and a CPU baseline each (sizes: see `scale`;
             --full runs the BASELINE sizes)
  strong     ONE archive of configs 3, 5 (full size) and 4 (half size) split over the N
             ranks by the library's partitioner (b2i_partition_contiguous / _lpt),
             device-resident; N = 1 is the base of the series
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified libarchive + zlib)
             on the box's host cores, bounded sample

Multi-GPU (--gpus N under torchrun): the headline is weak scaling - every rank decodes its
own archive of the config's shape, no data-path collective; torch.distributed (NCCL) only
carries the barrier and the max-over-ranks of the device time.

`--impl reference` times the reference CPU implementation (rank 0 only) and touches
nothing of this repo's library.
"""
import argparse
import ctypes as C
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

GIB = 1 << 30
CONFIGS = {
    # name: (BASELINE config number, container kind, description, default scale)
    "zip64k": (1, "zip", "ZIP of 4096 x 64 KiB synthetic-text entries, deflate level 6 (256 MiB out), CRC check on", 1.0),
    "stored1m": (2, "zip", "ZIP of 1024 x 1 MiB stored (method 0) entries: CRC-32 verification only", 1.0),
    "bgzf64k": (3, "bgzf", "BGZF multi-member gzip, 65536 members of 64 KiB (4 GiB out)", 0.25),
    "mixed": (4, "zip", "ZIP64 of 8 GiB, log-uniform entry sizes 1 KiB-16 MiB, dynamic/fixed/stored blocks mixed", 0.25),
    "tiny4k": (5, "zip", "ZIP64 of 500k x 4 KiB text entries (2 GB)", 0.25),
}


def load_synth():
    """libarchive_b200/synth.py by path: the generators are pure Python/numpy, and the
    reference arm must not import the package (which binds this repo's library)."""
    spec = importlib.util.spec_from_file_location("b2i_synth", os.path.join(ROOT, "libarchive_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2i_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


def shape(name, scale):
    """What both arms print as `config`: only things the workload definition fixes."""
    cfg, kind, desc, _ = CONFIGS[name]
    if name == "zip64k":
        units, out = max(8, int(4096 * scale)), None
        out = units * 65536
    elif name == "stored1m":
        units = max(2, int(1024 * scale)); out = units << 20
    elif name == "bgzf64k":
        units = max(8, int(65536 * scale)); out = units * 65280
    elif name == "mixed":
        units, out = None, max(8 << 20, int(8 * GIB * scale))
    else:
        units = max(64, int(500_000 * scale)); out = units * 4096
    # timing rule: inputs larger than L2 instead of a flush between iterations - every step
    # reads its compressed bytes and writes `out` bytes, more than the 126 MB L2 for every
    # workload at the default and full scales (the GPU arm checks it against the real sizes)
    return {"workload": name, "baseline_config": cfg, "description": desc, "scale": scale,
            "units": units, "out_bytes": out, "data": "synthetic, seeded",
            "l2": "no flush: each step touches its input + %.0f MB of output, L2 is 126 MB" % (out / 1e6)}


def build_workload(name, rank, scale=1.0):
    synth = load_synth()
    threads = max(4, min(32, (os.cpu_count() or 8)))
    sh = shape(name, scale)
    if name == "zip64k":
        n = sh["units"]
        parts = synth.split_text(n * 65536, 65536, 12345 + rank)
        return synth.make_zip([synth.ZipMember("e%06d.txt" % i, p) for i, p in enumerate(parts)],
                              threads=threads), "zip"
    if name == "stored1m":
        n = sh["units"]
        blob = synth.synth_random(n << 20, 2 + rank)
        return synth.make_zip([synth.ZipMember("s%05d.bin" % i, blob[i << 20:(i + 1) << 20], method=0)
                               for i in range(n)], threads=1), "zip"
    if name == "bgzf64k":
        n = sh["units"]
        parts = synth.text_corpus_parts(n * 65280, 65280, 54321 + rank)[:n]
        return synth.make_bgzf(parts, threads=threads), "bgzf"
    if name == "tiny4k":
        n = sh["units"]
        parts = synth.text_corpus_parts(n * 4096, 4096, 5 + rank)[:n]
        return synth.make_zip([synth.ZipMember("t%06d" % i, p) for i, p in enumerate(parts)],
                              zip64=True, threads=threads), "zip"
    if name == "mixed":
        return synth.config4_zip64_mixed(total=sh["out_bytes"], seed=4 + rank, threads=threads), "zip"
    raise SystemExit("unknown workload " + name)


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

def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(archive, kind, steps, warmup, sample_units=None, procs=None, budget_s=None):
    """Time the reference's own CPU implementation (oracle/_ref = unmodified libarchive +
    zlib; the scalar oracle port when _ref is absent) on a bounded sample.  Uses nothing
    of libarchive_b200: oracle_extract counts the entries itself."""
    nproc = procs or usable_cores()
    extract = os.path.join(ROOT, "oracle", "_ref", "oracle_extract")
    if os.path.exists(extract):
        with tempfile.NamedTemporaryFile(prefix="b2i_cpu_", suffix=".bin", delete=False, dir="/tmp") as f:
            f.write(archive)
            path = f.name
        try:
            cmd = [extract, "bench", path, "--procs", str(nproc), "--reps", "3"]
            if sample_units:
                cmd += ["--limit", str(sample_units)]
            if kind != "zip":
                cmd.append("--raw")
            times, out_bytes, j = [], 0, {}
            t_start = time.perf_counter()
            for i in range(warmup + steps):
                r = subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=300)
                j = json.loads(r.stdout.strip().splitlines()[-1])
                if j.get("bad"):
                    raise RuntimeError("reference reported errors: %s" % j)
                if i >= warmup:
                    times.append(j["seconds"])
                out_bytes = j["out_bytes"]
                if budget_s and times and time.perf_counter() - t_start > budget_s:
                    break
        finally:
            os.unlink(path)
        sec = sum(times) / len(times)
        return {"value": out_bytes / sec / 1e9, "unit": "GB/s", "cores": nproc, "kind": "reference",
                "sample": "%s of the workload's %s via oracle/_ref (unmodified libarchive + zlib %s), one process "
                          "per core on %d usable cores, archive_read_data into a 64 KiB buffer, CRC on, best of 3 "
                          "per step" % (j.get("units", "?"), "entries" if kind == "zip" else "members",
                                        j.get("zlib", "?"), nproc),
                "seconds": sec, "out_bytes": out_bytes}
    # oracle port (scalar, 1 thread): only when the reference could not be built
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, ROOT)
    import oracle_binding as ob
    from libarchive_b200 import capi, reader
    entries, _, _ = capi.zip_index(archive)
    descs, out_bytes, _ = reader.plan_zip(entries)
    units = min(len(descs), 64)
    sub = (type(descs[0]) * units)(*descs[:units])
    need = max(int(d.out_off + d.out_cap) for d in sub)
    t0 = time.perf_counter()
    res, _ = ob.decode_batch(archive, sub, need)
    sec = time.perf_counter() - t0
    nbytes = sum(int(r.out_bytes) for r in res)
    return {"value": nbytes / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
            "sample": "%d streams through the scalar oracle port" % units, "seconds": sec, "out_bytes": nbytes}


def cpu_sample_units(name, scale):
    """Bounded sample (about 10 s of CPU work on 16 cores at ~0.25 GB/s per core)."""
    sh = shape(name, scale)
    if name == "zip64k":
        return sh["units"]
    if name == "stored1m":
        return sh["units"]
    if name == "bgzf64k":
        return min(sh["units"], 16384)
    if name == "tiny4k":
        return min(sh["units"], 131072)
    return 300          # mixed: the first 300 entries (about 0.5 GB)


# --------------------------------------------------------------------------- GPU arm helpers

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


class Rig:
    """One rank's device context, stream and torch handles."""

    def __init__(self, local_rank, world):
        import torch
        import torch.distributed as dist
        from libarchive_b200 import capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.world, self.local_rank = world, local_rank
        torch.cuda.set_device(local_rank)
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a dedicated stream shared by torch (events) and the library (kernels, copies);
        # the legacy default stream has handle 0, which the C ABI reads as "make your own"
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.ctx = capi.Context(local_rank, self.stream.cuda_stream)
        self.L = capi.lib()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)

    def reduce(self, values, op="max"):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t]


def device_resident(rig, archive_addr, in_bytes, descs, out_bytes, K, W):
    """K timed launches of one plan over device-resident input; -> (per-step ms, total ms, launches)."""
    L, ctx, capi = rig.L, rig.ctx, rig.capi
    n = len(descs)
    d_in = L.b2i_device_alloc(ctx.h, in_bytes + 64)
    d_out = L.b2i_device_alloc(ctx.h, out_bytes + 64)
    assert d_in and d_out
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, archive_addr, in_bytes))
    plan = C.c_void_p()
    ctx._check(L.b2i_plan_create(ctx.h, descs, n, C.byref(plan)))
    res = (capi.StreamResult * max(n, 1))()

    def step():
        ctx._check(L.b2i_plan_launch(plan, d_in, in_bytes, d_out, out_bytes))

    # correctness gate before timing: every stream OK, sizes and CRCs as the directory says
    step()
    ctx._check(L.b2i_plan_results(plan, res))
    bad = [i for i in range(n) if res[i].status != 0 or res[i].flags != 0]
    if bad:
        raise SystemExit("bench.py: %d streams failed verification (first: %d status %d flags %d)" %
                         (len(bad), bad[0], res[bad[0]].status, res[bad[0]].flags))
    for _ in range(W):
        step()
    rig.barrier()
    launches0 = ctx.launch_count
    evs = [rig.event() for _ in range(K + 1)]
    evs[0].record(rig.stream)
    for i in range(K):
        step()
        evs[i + 1].record(rig.stream)
    rig.barrier()
    launches = ctx.launch_count - launches0
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    total_ms = evs[0].elapsed_time(evs[K])
    L.b2i_plan_destroy(plan)
    return step_ms, total_ms, launches, d_in, d_out


def link_bandwidth(rig, d_buf, h_buf, nbytes, reps=4):
    """Pinned H2D / D2H GB/s of this rank; with world > 1 all ranks copy at the same time."""
    L, ctx = rig.L, rig.ctx
    out = []
    for fn in (lambda: ctx._check(L.b2i_memcpy_h2d(ctx.h, d_buf, h_buf, nbytes)),
               lambda: ctx._check(L.b2i_memcpy_d2h(ctx.h, h_buf, d_buf, nbytes))):
        best = 0.0
        for _ in range(reps):
            rig.barrier()
            a, b = rig.event(), rig.event()
            a.record(rig.stream)
            fn()
            b.record(rig.stream)
            rig.torch.cuda.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
        out.append(best)
    return out


def abi_e2e(rig, h_in, in_bytes, descs, out_bytes, K, W):
    """C-ABI end to end with pinned host buffers, copies inside the timed region:
    -> (ms per step one call at a time, ms per step two jobs in flight, ms per step b2i_pipe)."""
    L, ctx, capi = rig.L, rig.ctx, rig.capi
    n = len(descs)
    h_out = L.b2i_host_alloc(out_bytes + 64)
    h_out2 = L.b2i_host_alloc(out_bytes + 64)
    assert h_out and h_out2
    res = (capi.StreamResult * max(n, 1))()

    def check(r):
        bad = [i for i in range(n) if r[i].status != 0 or r[i].flags != 0]
        if bad:
            raise SystemExit("bench.py: e2e pass failed verification for %d streams" % len(bad))

    for _ in range(min(W, 2)):
        ctx._check(L.b2i_decode_host(ctx.h, h_in, in_bytes, descs, n, h_out, out_bytes, res))
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        ctx._check(L.b2i_decode_host(ctx.h, h_in, in_bytes, descs, n, h_out, out_bytes, res))
    rig.torch.cuda.synchronize()
    one_ms = (time.perf_counter() - t0) * 1e3 / K
    check(res)

    def pipelined(steps):
        jobs, outs = [], []
        for i in range(steps):
            if len(jobs) == 2:
                outs.append(ctx.wait(jobs.pop(0)))
            jobs.append(ctx.submit(h_in, in_bytes, descs, h_out if i % 2 == 0 else h_out2, out_bytes))
        while jobs:
            outs.append(ctx.wait(jobs.pop(0)))
        return outs

    pipelined(min(W, 2))
    rig.barrier()
    t0 = time.perf_counter()
    outs = pipelined(K)
    rig.torch.cuda.synchronize()
    two_ms = (time.perf_counter() - t0) * 1e3 / K
    for r in outs[-2:]:
        check(r)

    # the streaming engine the plugins use: windows through a ring of pinned buffers
    def piped():
        p = capi.Pipe([ctx], descs, mem=h_in, mem_size=in_bytes)
        try:
            for i in range(0, n, 64):
                p.get(min(i + 63, n - 1))
                p.release(i)
            _, _, r = p.get(n - 1)
            if r.status != 0 or r.flags != 0:
                raise SystemExit("bench.py: pipe pass failed verification")
        finally:
            p.close()

    for _ in range(min(W, 2)):
        piped()
    rig.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        piped()
    pipe_ms = (time.perf_counter() - t0) * 1e3 / K
    L.b2i_host_free(h_out)
    L.b2i_host_free(h_out2)
    return one_ms, two_ms, pipe_ms


def public_api_e2e(archive, kind, local_rank, K, W):
    """libarchive's public API on the drop-in library, in a separate process (api_bench):
    archive image in pageable memory -> archive_read_data_block / archive_read_data."""
    exe = os.path.join(ROOT, "libarchive_b200", "api_bench")
    if not os.path.exists(exe):
        return None
    with tempfile.NamedTemporaryFile(prefix="b2i_api_", suffix=".bin", delete=False, dir="/dev/shm"
                                     if os.path.isdir("/dev/shm") else "/tmp") as f:
        f.write(archive)
        path = f.name
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = str(local_rank) if "CUDA_VISIBLE_DEVICES" not in os.environ else \
        os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]
    out = {}
    try:
        for mode in ("block", "data"):
            cmd = [exe, path, "--mode", mode, "--steps", str(K), "--warmup", str(max(2, min(W, 3)))]
            if kind != "zip":
                cmd.append("--raw")
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
            if r.returncode != 0:
                out[mode] = {"error": (r.stderr or r.stdout).strip()[-300:]}
                continue
            out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    finally:
        os.unlink(path)
    return out


def measure_config(rig, name, scale, rank, K, W, peak, with_cpu, seed_rank=None):
    """One BASELINE config on this rank's GPU: device-resident + C-ABI e2e (+ CPU sample)."""
    capi, L, ctx = rig.capi, rig.L, rig.ctx
    t_gen = time.perf_counter()
    archive, kind = build_workload(name, rank if seed_rank is None else seed_rank, scale)
    gen_s = time.perf_counter() - t_gen
    descs, out_bytes, usize, csize = plan_for(archive, kind)
    in_bytes = len(archive)
    h_in = L.b2i_host_alloc(in_bytes + 64)
    assert h_in
    C.memmove(h_in, archive, in_bytes)
    note("  %s: generated in %.1fs, %d streams, device-resident" % (name, gen_s, len(descs)))
    step_ms, total_ms, launches, d_in, d_out = device_resident(rig, h_in, in_bytes, descs, max(out_bytes, 16), K, W)
    one_ms = two_ms = pipe_ms = None
    if out_bytes:
        note("  %s: C-ABI e2e" % name)
        one_ms, two_ms, pipe_ms = abi_e2e(rig, h_in, in_bytes, descs, out_bytes, max(2, K // 2), W)
    L.b2i_device_free(ctx.h, d_in)
    L.b2i_device_free(ctx.h, d_out)
    L.b2i_host_free(h_in)
    tot_ms, = rig.reduce([total_ms], "max")
    usz, csz = rig.reduce([float(usize), float(csize)], "sum")
    kern_ms = statistics.mean(step_ms)
    alg = algorithmic_bytes(descs)
    crc_only = name == "stored1m"
    work = csz if crc_only else usz
    out = {
        "baseline_config": CONFIGS[name][0], "metric": "crc32_GBps" if crc_only else "inflate_out_GBps",
        "value": work * K / (tot_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": tot_ms / K,
        "config": shape(name, scale), "streams_per_gpu": len(descs), "out_bytes_per_gpu": usize,
        "in_bytes_per_gpu": csize, "gpu_launches": int(launches), "generate_s": round(gen_s, 1),
        "roofline": {"bound": "hbm", "kernel": "b2i_crc_chunks_kernel" if crc_only else "b2i_inflate_kernel(+team)",
                     "achieved": alg / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (kern_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg,
                     "launch_ms": kern_ms, "out_frac_of_hbm": (usize / (kern_ms * 1e-3) / 1e9) / peak},
    }
    if one_ms is not None:
        ms = rig.reduce([one_ms, two_ms, pipe_ms], "max")
        out["e2e"] = {"unit": "GB/s", "h2d_bytes_per_step": csize + len(descs) * 48,
                      "d2h_bytes_per_step": out_bytes + len(descs) * 32,
                      "b2i_decode_host": usz / (ms[0] * 1e-3) / 1e9,
                      "b2i_submit_wait_two_jobs": usz / (ms[1] * 1e-3) / 1e9,
                      "b2i_pipe": usz / (ms[2] * 1e-3) / 1e9,
                      "value": usz / (min(ms) * 1e-3) / 1e9}
    if with_cpu:
        try:
            note("  %s: cpu baseline" % name)
            cb = cpu_reference_run(archive, kind, 1, 0, sample_units=cpu_sample_units(name, scale))
            out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the baseline must not sink the GPU number
            out["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": str(ex)}
    return out, archive, kind


def strong_scaling(rig, name, scale, rank, K, W):
    """ONE archive split over the ranks by the library's partitioner; device-resident."""
    from libarchive_b200 import shard
    capi, L, ctx = rig.capi, rig.L, rig.ctx
    if rig.world == 1:
        archive, kind = build_workload(name, 0, scale)
    else:
        # rank 0 builds the archive once; the others read it from shared memory
        path = "/dev/shm/b2i_strong_%s_%d.bin" % (name, os.getppid())
        alt = "/tmp/b2i_strong_%s_%d.bin" % (name, os.getppid())
        if rank == 0:
            archive, kind = build_workload(name, 0, scale)
            try:
                with open(path + ".tmp", "wb") as f:
                    f.write(archive)
                os.rename(path + ".tmp", path)
            except OSError:                      # a small /dev/shm: the file system will do
                for pth in (path + ".tmp", path):
                    if os.path.exists(pth):
                        os.unlink(pth)
                with open(alt + ".tmp", "wb") as f:
                    f.write(archive)
                os.rename(alt + ".tmp", alt)
        rig.barrier()
        if not os.path.exists(path):
            path = alt
        if rank != 0:
            with open(path, "rb") as f:
                archive = f.read()
            kind = CONFIGS[name][1]
        rig.barrier()
        if rank == 0:
            os.unlink(path)
    descs, out_bytes, usize, csize = plan_for(archive, kind)
    lpt = name == "mixed"
    if lpt:
        sub, in_lo, in_hi, sub_out, idx = shard.shard_descs_lpt(descs, rank, rig.world)
        # an LPT share is scattered over the archive: keep archive offsets, stage the whole image
        sub, in_lo, in_hi = capi.make_descs([capi.StreamDesc.from_buffer_copy(descs[i]) for i in idx]), 0, len(archive)
        o = 0
        for d in sub:
            d.out_off = o
            o = (o + int(d.out_cap) + 15) & ~15
        sub_out = o
    else:
        sub, in_lo, in_hi, sub_out, _ = shard.shard_descs(descs, rank, rig.world)
    span = in_hi - in_lo
    h_in = L.b2i_host_alloc(span + 64)
    if span:
        C.memmove(h_in, archive[in_lo:in_hi], span)
    my_out = sum(int(d.expect_out) for d in sub)
    step_ms, total_ms, launches, d_in, d_out = device_resident(rig, h_in, span, sub, max(sub_out, 16), K, W)
    L.b2i_device_free(ctx.h, d_in)
    L.b2i_device_free(ctx.h, d_out)
    L.b2i_host_free(h_in)
    tot_ms, = rig.reduce([total_ms], "max")
    mn_ms = -rig.reduce([-total_ms], "max")[0]
    biggest = max((int(d.in_len) + int(d.out_cap) for d in descs), default=0)
    return {"baseline_config": CONFIGS[name][0], "value": usize * K / (tot_ms * 1e-3) / 1e9, "unit": "GB/s",
            "ms_per_step": tot_ms / K, "fastest_rank_ms_per_step": mn_ms / K, "scaling": "strong",
            "partition": "b2i_partition_lpt" if lpt else "b2i_partition_contiguous",
            "config": shape(name, scale), "streams": len(descs), "out_bytes": usize,
            "largest_stream_share": biggest / max(1, usize + csize),
            "this_rank_streams": len(sub), "this_rank_out_bytes": my_out}


# --------------------------------------------------------------------------- main

def note(msg):
    """Progress on stderr (the driver keeps stdout for the one JSON line)."""
    if os.environ.get("B2I_BENCH_QUIET") is None:
        sys.stderr.write("[bench %7.1fs rank %s] %s\n" % (time.perf_counter() - T_START, os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


T_START = time.perf_counter()


def main():
    import faulthandler
    # a hang must not eat the box: dump every thread's stack and leave
    faulthandler.dump_traceback_later(int(os.environ.get("B2I_BENCH_WATCHDOG", "1500")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="zip64k", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=None, help="size of the headline workload relative to BASELINE")
    ap.add_argument("--full", action="store_true", help="configs 3-5 at the BASELINE sizes (4 GiB / 8 GiB / 500k)")
    ap.add_argument("--only", action="store_true", help="just the headline workload: no configs / strong objects")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = args.steps
    head = args.workload
    head_scale = args.scale if args.scale is not None else (1.0 if args.full else CONFIGS[head][3])

    def scale_of(name):
        return 1.0 if args.full else CONFIGS[name][3]

    if args.impl == "reference":
        if rank != 0:
            return 0
        archive, kind = build_workload(head, 0, head_scale)
        cb = cpu_reference_run(archive, kind, K, W, sample_units=cpu_sample_units(head, head_scale), budget_s=150)
        line = {"metric": "inflate_out_GBps" if head != "stored1m" else "crc32_GBps", "value": cb["value"],
                "unit": "GB/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": cb["seconds"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "impl": "reference", "config": shape(head, head_scale),
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    sys.path.insert(0, ROOT)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    # one process per GPU: staging buffers on the GPU's own NUMA node (multi-socket boxes)
    from libarchive_b200.shard import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 and not os.environ.get("B2I_NO_NUMA_BIND") else 0
    rig = Rig(local_rank, world)
    capi, L, ctx = rig.capi, rig.L, rig.ctx
    peak, peak_src = measured_hbm_peak()

    # ---- headline ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    with_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    note("headline %s" % head)
    hd, archive, kind = measure_config(rig, head, head_scale, rank, K, W, peak, with_cpu)
    clocks = sampler.stop()
    note("headline done: %.1f GB/s device-resident" % hd["value"])

    # host link, this rank alone is meaningless at N > 1: all ranks copy at once
    nlink = 256 << 20
    h_l = L.b2i_host_alloc(nlink)
    d_l = L.b2i_device_alloc(ctx.h, nlink)
    h2d, d2h = link_bandwidth(rig, d_l, h_l, nlink)
    L.b2i_host_free(h_l)
    L.b2i_device_free(ctx.h, d_l)
    h2d_sum, d2h_sum = rig.reduce([h2d, d2h], "sum")
    h2d_min, d2h_min = [-x for x in rig.reduce([-h2d, -d2h], "max")]

    # ---- e2e through libarchive's public API (drop-in library, separate process) --------
    rig.barrier()
    note("public API e2e")
    api = public_api_e2e(archive, kind, local_rank, max(3, K // 4), W) if head != "stored1m" else None
    note("public API e2e done: %s" % (api,))
    usz_all, = rig.reduce([float(hd["out_bytes_per_gpu"])], "sum")
    e2e = dict(hd.get("e2e", {}))
    e2e["c_abi_best"] = e2e.pop("value", None)
    if api and "seconds_mean" in api.get("block", {}):
        blk_s, dat_s = rig.reduce([api["block"]["seconds_mean"],
                                   api.get("data", {}).get("seconds_mean", float("nan"))], "max")
        e2e["value"] = usz_all / blk_s / 1e9
        e2e["call"] = ("libarchive public API on libarchive_dropin.so, one consumer thread per GPU: "
                       "archive_read_open_memory (pageable image) + archive_read_next_header + "
                       "archive_read_data_block; decoded bytes are in host memory when the call returns")
        e2e["archive_read_data_64KiB"] = usz_all / dat_s / 1e9 if dat_s == dat_s else None
        e2e["archive_read_data_note"] = ("adds libarchive's own per-block memcpy into the caller's 64 KiB buffer "
                                         "(archive_read.c:879), one thread")
    else:
        e2e["value"] = e2e["c_abi_best"]
        e2e["call"] = "C ABI (api_bench unavailable: %s)" % (api or "not built")
    e2e["ms_per_step"] = usz_all / (e2e["value"] * 1e9) * 1e3 if e2e.get("value") else None
    per_gpu_out = hd["out_bytes_per_gpu"]
    link_floor_ms = max(hd["in_bytes_per_gpu"] / (h2d_min * 1e9), per_gpu_out / (d2h_min * 1e9)) * 1e3
    e2e["host_link"] = {"h2d_GBps_per_gpu_concurrent": h2d_min, "d2h_GBps_per_gpu_concurrent": d2h_min,
                        "h2d_GBps_all_gpus": h2d_sum, "d2h_GBps_all_gpus": d2h_sum,
                        "frac_of_link": link_floor_ms / e2e["ms_per_step"] if e2e.get("ms_per_step") else None,
                        "note": "pinned cudaMemcpyAsync, all ranks copying at the same time; frac = time the slower "
                                "direction needs at that bandwidth / measured e2e step time"}

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % head)
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roof = hd["roofline"]
    roof.update({"traffic": traffic, "peak_source": peak_src})
    if head != "stored1m":
        roof["kernel"] = "b2i_inflate_kernel"

    line = {
        "metric": hd["metric"], "value": hd["value"], "unit": "GB/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": hd["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": shape(head, head_scale),
        "workload_detail": {"streams_per_gpu": hd["streams_per_gpu"], "out_bytes_per_gpu": hd["out_bytes_per_gpu"],
                            "in_bytes_per_gpu": hd["in_bytes_per_gpu"],
                            "l2": "working set (in+out %.0f MB) exceeds the 126 MB L2; no flush" %
                                  ((hd["out_bytes_per_gpu"] + hd["in_bytes_per_gpu"]) / 1e6),
                            "parallelism": "every rank its own archive, no collective", "numa_bound_cpus": numa_cpus},
        "roofline": roof,
        "e2e": e2e,
        "gpu_launches": hd["gpu_launches"],
        "clocks": clocks,
    }
    if "cpu_baseline" in hd:
        line["cpu_baseline"] = hd["cpu_baseline"]
        if hd["cpu_baseline"].get("value"):
            try:
                one = cpu_reference_run(archive, kind, 1, 0, sample_units=512, procs=1)
                line["cpu_baseline"]["one_core_value"] = one["value"]
            except Exception:
                pass
    elif rank == 0:
        line["cpu_baseline"] = None
    del archive

    # ---- the other BASELINE configs --------------------------------------------------
    if not args.only:
        Kc, Wc = max(3, min(K, 5)), 3
        configs = {}
        for name in CONFIGS:
            if name == head:
                continue
            try:
                note("config %s" % name)
                c, a_, _ = measure_config(rig, name, scale_of(name), rank, Kc, Wc, peak, with_cpu)
                del a_
                c["steps"], c["warmup"] = Kc, Wc
                configs[name] = c
            except SystemExit as ex:
                configs[name] = {"error": str(ex)}
        line["configs"] = configs
        # ONE archive split over the N ranks (N = 1: the whole of it, the base of the series);
        # large enough that a rank's share still fills its GPU at N = 8
        strong = {}
        for name, sc in (() if args.full and world == 1 else (("bgzf64k", 1.0), ("tiny4k", 1.0), ("mixed", 0.5))):
            note("strong %s" % name)
            try:
                strong[name] = strong_scaling(rig, name, sc, rank, Kc, Wc)
            except SystemExit as ex:
                strong[name] = {"error": str(ex)}
        line["strong"] = strong

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        rig.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
