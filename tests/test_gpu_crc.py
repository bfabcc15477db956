"""GPU parity for CRC-32 (K2/K3): scalar drop-in contract of archive_crc32.h:43-84
and stored ZIP entries (CRC in place / copy + CRC)."""
import ctypes as C
import zlib

import numpy as np
import pytest

import oracle_binding as ob
from libarchive_b200 import capi, synth
from libarchive_b200.capi import StreamDesc

pytestmark = pytest.mark.gpu


def test_crc32_contract(ctx):
    assert ctx.crc32(None) == 0                       # crc32(x, NULL, 0) == 0
    assert ctx.crc32(b"") == 0
    assert ctx.crc32(b"", 0x1234) == 0x1234
    assert ctx.crc32(b"hello\nhello\nhello\n") == 0x3D66373A   # test_read_format_zip.zip file1
    assert ctx.crc32(b"123456789") == 0xCBF43926


def test_crc32_lengths_alignments_and_chaining(ctx):
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, (1 << 20) + 64, dtype=np.uint8).tobytes()
    for n in [1, 2, 3, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1023, 4095,
              4096, 4097, 16383, 16384, 16385, 65535, 65536, 65537, 100003, (1 << 20) + 3]:
        for off in (0, 1, 5):
            piece = data[off:off + n]
            assert ctx.crc32(piece) == ob.crc32(piece) == (zlib.crc32(piece) & 0xFFFFFFFF), (n, off)
    # incremental chaining (test_write_format_zip_compression_store.c:137-138 pins it)
    c = 0
    ref = 0
    for lo, hi in [(0, 10), (10, 4000), (4000, 70000), (70000, 1 << 20)]:
        c = ctx.crc32(data[lo:hi], c)
        ref = zlib.crc32(data[lo:hi], ref)
    assert c == ref & 0xFFFFFFFF


def test_crc32_combine_matches_direct():
    rng = np.random.default_rng(2)
    L = capi.lib()
    for la, lb in [(0, 0), (0, 5), (5, 0), (1, 1), (17, 3), (1000, 1), (4096, 65536), (99999, 123457)]:
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        got = L.b2i_crc32_combine(zlib.crc32(a), zlib.crc32(b), lb)
        assert got == (zlib.crc32(a + b) & 0xFFFFFFFF) == ob.lib().orc_crc32_combine(zlib.crc32(a), zlib.crc32(b), lb)


@pytest.mark.parametrize("no_copy", [True, False])
def test_stored_entries(ctx, no_copy):
    rng = np.random.default_rng(3)
    sizes = [0, 1, 15, 16, 17, 1000, 16384, 16385, 50000, 1 << 20, (1 << 20) + 7, 3 * (1 << 20) + 11]
    blob = bytearray()
    items, out = [], 0
    for k, n in enumerate(sizes):
        blob += b"\x00" * (k % 5)                     # arbitrary alignment of each body
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        d = StreamDesc()
        d.in_off, d.in_len, d.method = len(blob), n, 0
        d.expect_out = n
        d.expect_crc = zlib.crc32(data) & 0xFFFFFFFF if k != 4 else 0xDEADBEEF
        d.flags = capi.F_NO_COPY if no_copy else 0
        d.out_off, d.out_cap = out, 0 if no_copy else n
        if not no_copy:
            out = (out + n + 15) & ~15
        blob += data
        items.append(d)
    descs = capi.make_descs(items)
    blob = bytes(blob)
    inbuf = C.create_string_buffer(blob, len(blob) + 32)
    outbuf = C.create_string_buffer(out + 32)
    res = ctx.decode_host(inbuf, len(blob), descs, outbuf, out)
    ores, oout = ob.decode_batch(blob, descs, out)
    for k, n in enumerate(sizes):
        assert (res[k].status, res[k].crc, res[k].out_bytes, res[k].in_bytes, res[k].flags) == \
               (ores[k].status, ores[k].crc, ores[k].out_bytes, ores[k].in_bytes, ores[k].flags), (k, n)
    assert res[4].flags & capi.R_CRC_MISMATCH
    if not no_copy:        # compare entry by entry: alignment gaps are never written by either side
        for k, n in enumerate(sizes):
            o = descs[k].out_off
            assert outbuf.raw[o:o + n] == oout[o:o + n], (k, n)


def test_unsupported_method(ctx):
    d = StreamDesc()
    d.in_off, d.in_len, d.method = 0, 10, 12
    inbuf = C.create_string_buffer(64)
    res = ctx.decode_host(inbuf, 32, capi.make_descs([d]), None, 0)
    assert res[0].status == capi.S_UNSUPPORTED
