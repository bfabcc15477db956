import sys, time, zlib, os, ctypes as C
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
from libarchive_b200.capi import StreamDesc
ctx = capi.Context(0)
L = capi.lib()
txt = synth.synth_text(8 << 20, 3)
rnd = synth.synth_random(8 << 20, 4)
cases = [("text-dynamic", synth.deflate_raw(txt, 6)), ("text-fixed", synth.deflate_raw(txt, 1, zlib.Z_FIXED)),
         ("random-stored", synth.deflate_raw(rnd, 6))]
for name, s in cases:
    d = StreamDesc(); d.in_off = 0; d.in_len = len(s); d.out_cap = 8 << 20; d.expect_out = 8 << 20; d.method = 8
    d.expect_crc = zlib.crc32(txt if "text" in name else rnd) & 0xFFFFFFFF
    descs = capi.make_descs([d])
    d_in = L.b2i_device_alloc(ctx.h, len(s) + 64); d_out = L.b2i_device_alloc(ctx.h, (8 << 20) + 64)
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, s, len(s)))
    plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, descs, 1, C.byref(plan)))
    res = (capi.StreamResult * 1)()
    for it in range(3):
        ctx.sync(); t = time.perf_counter()
        ctx._check(L.b2i_plan_launch(plan, d_in, len(s), d_out, 8 << 20)); ctx._check(L.b2i_plan_results(plan, res))
        ms = (time.perf_counter() - t) * 1e3
    print(os.environ.get("B2I_TEAM_MIN_BYTES", "team"), name, len(s), "ms %.1f" % ms, "MB/s %.0f" % ((8 << 20) / ms / 1e3), "status", res[0].status, res[0].flags)
    L.b2i_plan_destroy(plan)
