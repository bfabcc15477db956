"""Fixed cost of one host-buffer decode call (b2i_decode_host) for tiny batches."""
import ctypes as C, os, sys, time, zlib
sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, synth
from libarchive_b200.capi import StreamDesc
ctx = capi.Context(0)
L = capi.lib()
for usize, pinned in ((300, True), (300, False), (262150, True), (262150, False)):
    data = synth.synth_text(usize, 3) if usize < 100000 else bytes(usize)
    s = synth.deflate_raw(data, 6)
    d = StreamDesc(); d.in_len = len(s); d.out_cap = usize; d.expect_out = usize; d.method = 8
    d.expect_crc = zlib.crc32(data) & 0xFFFFFFFF
    descs = capi.make_descs([d])
    if pinned:
        h_in = L.b2i_host_alloc(len(s) + 64); h_out = L.b2i_host_alloc(usize + 64); C.memmove(h_in, s, len(s))
    else:
        bi = C.create_string_buffer(s, len(s) + 64); bo = C.create_string_buffer(usize + 64)
        h_in, h_out = C.addressof(bi), C.addressof(bo)
    res = (capi.StreamResult * 1)()
    for _ in range(20):
        ctx._check(L.b2i_decode_host(ctx.h, h_in, len(s), descs, 1, h_out, usize, res))
    t = time.perf_counter()
    N = 300
    for _ in range(N):
        ctx._check(L.b2i_decode_host(ctx.h, h_in, len(s), descs, 1, h_out, usize, res))
    us = (time.perf_counter() - t) / N * 1e6
    print("usize %7d csize %6d %s: %.0f us per call, status %d flags %d" % (usize, len(s), "pinned  " if pinned else "pageable", us, res[0].status, res[0].flags), flush=True)
