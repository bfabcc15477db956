"""One 16 MiB text stream through the team kernel, device-resident (for ncu)."""
import ctypes as C, os, sys, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
ctx = capi.Context(0)
L = capi.lib()
data = synth.synth_text(16 << 20, 3)
comp = synth.deflate_raw(data, 6)
d = capi.StreamDesc(); d.in_off, d.in_len, d.out_off, d.out_cap, d.expect_out = 0, len(comp), 0, len(data), len(data)
d.expect_crc, d.method = zlib.crc32(data), 8
descs = capi.make_descs([d])
d_in = L.b2i_device_alloc(ctx.h, len(comp) + 64); d_out = L.b2i_device_alloc(ctx.h, len(data) + 64)
buf = C.create_string_buffer(comp, len(comp) + 32)
ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, buf, len(comp)))
plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, descs, 1, C.byref(plan)))
res = (capi.StreamResult * 1)()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx._check(L.b2i_plan_launch(plan, d_in, len(comp), d_out, len(data)))
    ctx._check(L.b2i_plan_results(plan, res))
    assert res[0].status == 0 and res[0].flags == 0
print("ok")
