"""Per-phase cycle counters of the inflate kernel (profiling build: make -C libarchive_b200/csrc prof).
Run with B2I_LIB=libarchive_b200/libb200inflate_prof.so; b2i_ctx_sync prints and clears the counters."""
import sys, time, zlib, os, ctypes as C
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
from libarchive_b200.capi import StreamDesc

ctx = capi.Context(0)
L = capi.lib()


def run(name, streams, usizes, reps=2):
    descs, blob, off, ooff = [], bytearray(), 0, 0
    for s, u in zip(streams, usizes):
        d = StreamDesc(); d.in_off = off; d.in_len = len(s); d.out_off = ooff; d.out_cap = u; d.expect_out = u
        d.method = 8; d.flags = 2
        descs.append(d); blob += s; off += len(s); ooff += (u + 15) & ~15
    arr = capi.make_descs(descs)
    d_in = L.b2i_device_alloc(ctx.h, len(blob) + 64); d_out = L.b2i_device_alloc(ctx.h, ooff + 64)
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, bytes(blob), len(blob)))
    plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, arr, len(descs), C.byref(plan)))
    res = (capi.StreamResult * len(descs))()
    for it in range(reps):
        ctx.sync(); sys.stdout.flush(); t = time.perf_counter()
        ctx._check(L.b2i_plan_launch(plan, d_in, len(blob), d_out, ooff)); ctx._check(L.b2i_plan_results(plan, res))
        ms = (time.perf_counter() - t) * 1e3
        print(f"== {name} n={len(descs)} out={ooff} ms={ms:.2f} GB/s={ooff / ms / 1e6:.2f} status0={res[0].status}", flush=True)
    ctx.sync()
    L.b2i_plan_destroy(plan); L.b2i_device_free(ctx.h, d_in); L.b2i_device_free(ctx.h, d_out)


txt = synth.synth_text(8 << 20, 3)
one = synth.deflate_raw(txt, 6)
run("single-8MiB-text", [one], [8 << 20])
e = [synth.deflate_raw(txt[i * 65536:(i + 1) * 65536], 6) for i in range(128)]
run("one-64KiB", e[:1], [65536])
run("148x64KiB", [e[i % 128] for i in range(148)], [65536] * 148)
run("592x64KiB", [e[i % 128] for i in range(592)], [65536] * 592)
run("4096x64KiB", [e[i % 128] for i in range(4096)], [65536] * 4096)
