"""-m gpu parity on the SHAPES of BASELINE configs 3, 4 and 5 (config 1 at full size is in
test_gpu_reader.py, config 2 in test_gpu_crc.py): every stream decoded through the C ABI
is compared with the oracle - status, CRC, consumed / produced counts and bytes."""
import ctypes as C
import zlib
from concurrent.futures import ThreadPoolExecutor

import pytest

import oracle_binding as ob
from libarchive_b200 import capi, reader, synth

pytestmark = pytest.mark.gpu


def oracle_compare(archive: bytes, descs, gpu_res, gpu_out, threads: int = 16):
    """Entry by entry against the oracle; the oracle runs in `threads` slices (ctypes
    releases the GIL), each on descriptors rebased to its own output buffer."""
    n = len(descs)
    step = max(1, (n + threads - 1) // threads)

    def work(lo):
        hi = min(n, lo + step)
        sub = (capi.StreamDesc * (hi - lo))()
        base = int(descs[lo].out_off)
        need = 0
        for k in range(lo, hi):
            d = capi.StreamDesc.from_buffer_copy(descs[k])
            d.out_off = int(descs[k].out_off) - base
            need = max(need, int(d.out_off + d.out_cap))
            sub[k - lo] = d
        ores, oout = ob.decode_batch(archive, sub, need)
        bad = []
        for k in range(lo, hi):
            g, o = gpu_res[k], ores[k - lo]
            if (g.status, g.crc, g.out_bytes, g.in_bytes, g.flags) != (o.status, o.crc, o.out_bytes, o.in_bytes, o.flags):
                bad.append((k, "result", (g.status, g.crc, g.out_bytes, g.in_bytes, g.flags),
                            (o.status, o.crc, o.out_bytes, o.in_bytes, o.flags)))
                continue
            a, nb = int(descs[k].out_off), int(g.out_bytes)
            if gpu_out[a:a + nb] != oout[a - base:a - base + nb]:
                bad.append((k, "bytes"))
        return bad

    with ThreadPoolExecutor(threads) as ex:
        bad = [b for part in ex.map(work, range(0, n, step)) for b in part]
    assert not bad, bad[:5]


def gpu_decode(ctx, archive: bytes, descs, out_bytes: int):
    inbuf = C.create_string_buffer(archive, len(archive) + 32)
    outbuf = C.create_string_buffer(out_bytes + 32)
    res = ctx.decode_host(inbuf, len(archive), descs, outbuf, out_bytes)
    return res, memoryview(outbuf).cast("B")[:out_bytes]


def test_config4_mixed_1gib_with_16mib_entries(ctx):
    """ZIP64, log-uniform 1 KiB-16 MiB, dynamic / fixed / stored blocks mixed: 1 GiB of the
    shape plus forced 16 MiB entries of every kind (two dynamic, two fixed, one stored-block,
    two Z_FULL_FLUSH-interleaved)."""
    big = 16 << 20
    z = synth.config4_zip64_mixed(total=1 << 30, seed=44, threads=16,
                                  force=[(big, 0), (big, 0), (big, 3), (big, 3), (big, 6), (big, 9), (big - 12345, 9)])
    entries, _, _ = capi.zip_index(z)
    descs, out_bytes, which = reader.plan_zip(entries, stored_no_copy=False)
    assert len(descs) == len(entries) and sum(1 for d in descs if d.expect_out >= big - 12345) >= 7
    res, out = gpu_decode(ctx, z, descs, out_bytes)
    assert all(r.status == 0 and r.flags == 0 for r in res)
    oracle_compare(z, descs, res, out)


def test_config5_zip64_more_than_65535_tiny_entries(ctx):
    """> 65 535 entries of <= 4 KiB (ZIP64 end record), ragged sizes including 1-byte and
    empty-after-deflate payloads."""
    n = 70_001
    blob = synth.synth_text(16 << 20, 55)
    members, o = [], 0
    for i in range(n):
        s = 4096 if i % 3 == 0 else 1 + (i * 2654435761) % 4096
        if i % 1000 == 7:
            s = 1
        members.append(synth.ZipMember("t%06d" % i, blob[o % (len(blob) - 4096):][:s]))
        o += s
    z = synth.make_zip(members, zip64=True, threads=16)
    entries, _, _ = capi.zip_index(z)
    assert len(entries) == n
    descs, out_bytes, which = reader.plan_zip(entries)
    assert len(descs) == n
    res, out = gpu_decode(ctx, z, descs, out_bytes)
    assert all(r.status == 0 and r.flags == 0 for r in res)
    oracle_compare(z, descs, res, out)


def test_config3_bgzf_16384_members(ctx):
    """>= 16 384 BGZF members + the EOF member, one device pass, trailer CRC / ISIZE as
    expect_crc / expect_out."""
    parts = synth.text_corpus_parts(16384 * 65280, 65280, 333)[:16384]
    f = synth.make_bgzf(parts, threads=16)
    members, end = capi.gzip_scan_bgzf(f)
    assert len(members) == 16385 and end == len(f)
    descs, out_bytes = reader.plan_bgzf(members)
    res, out = gpu_decode(ctx, f, descs, out_bytes)
    assert all(r.status == 0 and r.flags == 0 for r in res)
    oracle_compare(f, descs, res, out)
    cc = 0
    for r in res[:16384]:
        cc = zlib.crc32(r.crc.to_bytes(4, "little"), cc)
    want = 0
    for p in parts:
        want = zlib.crc32((zlib.crc32(p) & 0xFFFFFFFF).to_bytes(4, "little"), want)
    assert cc == want


def test_stored_entry_larger_than_its_reserved_output_is_not_copied(ctx):
    """STORED, no NO_COPY, in_len > out_cap must not write past its reservation."""
    a, b = synth.synth_random(40000, 1), synth.synth_random(3000, 2)
    blob = a + b
    d0, d1 = capi.StreamDesc(), capi.StreamDesc()
    d0.in_off, d0.in_len, d0.out_off, d0.out_cap, d0.expect_out = 0, len(a), 0, 1024, len(a)
    d0.expect_crc, d0.method = zlib.crc32(a), 0
    d1.in_off, d1.in_len, d1.out_off, d1.out_cap, d1.expect_out = len(a), len(b), 1024, len(b), len(b)
    d1.expect_crc, d1.method = zlib.crc32(b), 0
    descs = capi.make_descs([d0, d1])
    total = 1024 + len(b) + 16
    inbuf = C.create_string_buffer(blob, len(blob) + 32)
    outbuf = C.create_string_buffer(b"\xAA" * (total + 64), total + 64)
    res = ctx.decode_host(inbuf, len(blob), descs, outbuf, total)
    assert res[0].status == capi.S_OUT_OVERFLOW
    assert res[1].status == 0 and res[1].flags == 0
    raw = outbuf.raw
    # entry 0's own 1024 reserved bytes are unspecified (the copy-out covers reservations);
    # what matters is that nothing was written past them
    assert raw[1024:1024 + len(b)] == b                 # entry 1 intact
    assert raw[total:total + 64] == b"\xAA" * 64           # nothing past the output buffer
