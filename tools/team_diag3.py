import ctypes as C, os, sys, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, reader, synth
total = 256 << 20
z = synth.config4_zip64_mixed(total=total, seed=4)
entries, _, _ = capi.zip_index(z)
descs, out_bytes, which = reader.plan_zip(entries, stored_no_copy=False)
ctx = capi.Context(0)
L = capi.lib()
inbuf = C.create_string_buffer(z, len(z) + 32)
outbuf = C.create_string_buffer(out_bytes + 32)
mode = sys.argv[1]
if mode == "host":
    res = ctx.decode_host(inbuf, len(z), descs, outbuf, out_bytes)
else:
    d_in = L.b2i_device_alloc(ctx.h, len(z) + 64); d_out = L.b2i_device_alloc(ctx.h, out_bytes + 64)
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, inbuf, len(z)))
    plan = C.c_void_p(); ctx._check(L.b2i_plan_create(ctx.h, descs, len(descs), C.byref(plan)))
    res = (capi.StreamResult * len(descs))()
    ctx._check(L.b2i_plan_launch(plan, d_in, len(z), d_out, out_bytes))
    ctx._check(L.b2i_plan_results(plan, res))
bad = [(k, res[k].status, res[k].flags, "%08x" % res[k].crc, "%08x" % descs[k].expect_crc, int(descs[k].expect_out)) for k in range(len(descs)) if res[k].status or res[k].flags]
print(mode, os.environ.get("B2I_PIPE_SLICES"), os.environ.get("B2I_TEAM_MIN_BYTES"), "bad:", bad)
