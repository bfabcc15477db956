"""BASELINE config 1 (4096 x 64 KiB text entries) through the single-warp kernel, device-resident,
CRC on: one launch per repetition (for ncu)."""
import ctypes as C, os, sys, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
ctx = capi.Context(0)
L = capi.lib()
n = 4096
parts = synth.split_text(n * 65536, 65536, 12345)
descs, blob, off, ooff = [], bytearray(), 0, 0
for p in parts:
    s = synth.deflate_raw(p, 6)
    d = capi.StreamDesc(); d.in_off, d.in_len, d.out_off, d.out_cap, d.expect_out = off, len(s), ooff, len(p), len(p)
    d.expect_crc, d.method = zlib.crc32(p), 8
    descs.append(d); blob += s; off += len(s); ooff += (len(p) + 15) & ~15
arr = capi.make_descs(descs)
d_in = L.b2i_device_alloc(ctx.h, len(blob) + 64); d_out = L.b2i_device_alloc(ctx.h, ooff + 64)
ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, bytes(blob), len(blob)))
plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, arr, n, C.byref(plan)))
res = (capi.StreamResult * n)()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx._check(L.b2i_plan_launch(plan, d_in, len(blob), d_out, ooff))
    ctx._check(L.b2i_plan_results(plan, res))
    assert all(r.status == 0 and r.flags == 0 for r in res)
print("ok in", len(blob), "out", ooff)
