"""GPU: one large stream alone, device-resident: MB/s per stream and (with the prof build,
B2I_LIB=.../libb200inflate_prof.so) the team kernel's phase clocks."""
import ctypes as C, os, sys, time, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
ctx = capi.Context(0)
L = capi.lib()
txt = synth.synth_text(16 << 20, 3)
cases = [("text-l6", txt, 6, 0), ("text-l1-fixed", txt, 1, zlib.Z_FIXED), ("text-l9", txt[:8 << 20], 9, 0),
         ("random-stored", synth.synth_random(16 << 20, 1), 6, 0)]
for name, data, level, strategy in cases:
    comp = synth.deflate_raw(data, level, strategy)
    d = capi.StreamDesc(); d.in_off, d.in_len, d.out_off, d.out_cap, d.expect_out = 0, len(comp), 0, len(data), len(data)
    d.expect_crc, d.method = zlib.crc32(data), 8
    descs = capi.make_descs([d])
    d_in = L.b2i_device_alloc(ctx.h, len(comp) + 64); d_out = L.b2i_device_alloc(ctx.h, len(data) + 64)
    buf = C.create_string_buffer(comp, len(comp) + 32)
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, buf, len(comp)))
    plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, descs, 1, C.byref(plan)))
    res = (capi.StreamResult * 1)()
    for rep in range(3):
        ctx.sync()
        t0 = time.perf_counter()
        ctx._check(L.b2i_plan_launch(plan, d_in, len(comp), d_out, len(data)))
        ctx._check(L.b2i_plan_results(plan, res))
        dt = time.perf_counter() - t0
    assert res[0].status == 0 and res[0].flags == 0, (res[0].status, res[0].flags)
    print("%-14s out %d in %d: %.2f ms = %.0f MB/s per stream" % (name, len(data), len(comp), dt * 1e3, len(data) / dt / 1e6), flush=True)
    ctx.sync()
    L.b2i_plan_destroy(plan); L.b2i_device_free(ctx.h, d_in); L.b2i_device_free(ctx.h, d_out)
