"""GPU diagnostic: decode a config-4 shaped archive and report, per failing stream, where the
decoded bytes first differ from zlib's."""
import ctypes as C
import os
import sys
import zlib

sys.path.insert(0, os.getcwd())
from libarchive_b200 import capi, reader, synth

total = int(sys.argv[1]) if len(sys.argv) > 1 else 256 << 20
z = synth.config4_zip64_mixed(total=total, seed=4)
entries, _, _ = capi.zip_index(z)
descs, out_bytes, which = reader.plan_zip(entries, stored_no_copy=False)
ctx = capi.Context(0)
inbuf = C.create_string_buffer(z, len(z) + 32)
outbuf = C.create_string_buffer(out_bytes + 32)
for rep in range(2):
    res = ctx.decode_host(inbuf, len(z), descs, outbuf, out_bytes)
    out = memoryview(outbuf).cast("B")
    nbad = 0
    for k, d in enumerate(descs):
        e = entries[which[k]]
        r = res[k]
        if r.status == 0 and r.flags == 0:
            continue
        nbad += 1
        comp = z[d.in_off:d.in_off + d.in_len]
        want = zlib.decompress(comp, -15)
        got = bytes(out[d.out_off:d.out_off + r.out_bytes])
        first = next((i for i in range(min(len(want), len(got))) if want[i] != got[i]), None)
        ndiff = sum(1 for i in range(0, min(len(want), len(got)), 1) if want[i] != got[i]) if first is not None else 0
        last = max((i for i in range(min(len(want), len(got))) if want[i] != got[i]), default=None)
        print("rep %d stream %d %s usize %d csize %d status %d flags %d out %d first_diff %s last_diff %s ndiff %d" %
              (rep, k, e["name"].decode(), len(want), d.in_len, r.status, r.flags, r.out_bytes, first, last, ndiff))
        if first is not None:
            print("   want", want[first - 8:first + 24])
            print("   got ", got[first - 8:first + 24])
    print("rep", rep, "bad", nbad, "of", len(descs))
