import ctypes as C, os, sys, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
ctx = capi.Context(0)
def one(name, data, level=6, strategy=0):
    comp = synth.deflate_raw(data, level, strategy)
    d = capi.StreamDesc(); d.in_off, d.in_len, d.out_off, d.out_cap, d.expect_out = 0, len(comp), 0, len(data), len(data)
    d.expect_crc, d.method = zlib.crc32(data), 8
    inbuf = C.create_string_buffer(comp, len(comp) + 32); outbuf = C.create_string_buffer(len(data) + 32)
    r = ctx.decode_host(inbuf, len(comp), capi.make_descs([d]), outbuf, len(data))[0]
    ok = outbuf.raw[:len(data)] == data
    print("%-10s n %9d status %d flags %d crc %08x want %08x bytes_ok %s" % (name, len(data), r.status, r.flags, r.crc, d.expect_crc, ok), flush=True)
for mb in (2.0, 7.9, 8.0, 8.1, 9.5, 15.9):
    one("rnd%.1f" % mb, synth.synth_random(int(mb * (1 << 20)), 5))
txt = synth.synth_text(16 << 20, 3)
for mb in (2.0, 8.1, 15.9):
    one("txt%.1f" % mb, txt[:int(mb * (1 << 20))])
one("zeros9", bytes(9 << 20))
