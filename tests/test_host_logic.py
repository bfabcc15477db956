"""CPU-only checks of the host side: the C ABI library loads and exports every symbol
include/b200inflate.h declares, the ZIP directory walk and the gzip/BGZF scan agree
with independent parsers, and the library fails loudly without a GPU."""
import io
import os
import re
import struct
import zipfile
import zlib

import pytest

from libarchive_b200 import capi, reader, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header():
    with open(os.path.join(ROOT, "include", "b200inflate.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(b2i_[a-z0-9_]+)\s*\(", hdr))
    L = capi.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
    assert declared == set(capi.EXPORTS)
    assert L.b2i_abi_version() == 4


def test_struct_layouts():
    import ctypes as C
    assert C.sizeof(capi.StreamDesc) == 48 and C.sizeof(capi.StreamResult) == 32


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.B2IError):
        capi.Context(0)


def test_zip_index_against_zipfile():
    parts = synth.split_text(50 * 5000, 5000, 3)
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w", zipfile.ZIP_DEFLATED) as z:
        for i, p in enumerate(parts):
            z.writestr("dir/f%03d.txt" % i, p)
        z.writestr("stored.bin", b"x" * 1000, zipfile.ZIP_STORED)
        z.writestr("empty", b"")
        z.writestr("d/", b"")
    blob = bio.getvalue()
    entries, corr, enc = capi.zip_index(blob)
    infos = sorted(zipfile.ZipFile(io.BytesIO(blob)).infolist(), key=lambda i: i.header_offset)
    assert corr == 0 and not enc and len(entries) == len(infos)
    for e, i in zip(entries, infos):
        assert e["name"] == i.filename.encode()
        assert (e["local_header_offset"], e["compressed_size"], e["uncompressed_size"], e["crc32"], e["method"]) == \
               (i.header_offset, i.compress_size, i.file_size, i.CRC, i.compress_type)
        if i.compress_size:
            raw = blob[e["data_offset"]:e["data_offset"] + e["compressed_size"]]
            data = zlib.decompress(raw, -15) if i.compress_type == 8 else raw
            assert zlib.crc32(data) == i.CRC
    assert (entries[-1]["mode"] & 0o170000) == 0o040000          # trailing slash => directory


@pytest.mark.parametrize("framing", ["sizes", "at_end"])
@pytest.mark.parametrize("zip64", [False, True])
def test_zip_index_framings_prefix_and_duplicates(framing, zip64):
    parts = synth.split_text(20 * 3000, 3000, 4)
    members = [synth.ZipMember("m%02d" % i, p) for i, p in enumerate(parts)]
    plain = synth.make_zip(members, framing=framing, zip64=zip64)
    e0, c0, _ = capi.zip_index(plain)
    # prepended data (SFX stub): the forward scan finds the directory, offsets are corrected
    padded = synth.make_zip(members, framing=framing, zip64=False, prefix=b"#!/bin/sh\n" * 100) if not zip64 else None
    assert len(e0) == 20 and c0 == 0
    for e, m in zip(e0, members):
        assert e["uncompressed_size"] == len(m.data) and e["crc32"] == zlib.crc32(m.data)
        assert zlib.decompress(plain[e["data_offset"]:e["data_offset"] + e["compressed_size"]], -15) == m.data
        assert not e["zip_flags"] & 8                                   # length-at-end cleared (zip.c:1108)
    if padded:
        e1, c1, _ = capi.zip_index(padded)
        assert c1 == 1000 and [x["data_offset"] - 1000 for x in e1] == [x["data_offset"] for x in e0]


def test_zip_index_many_entries_zip64_eocd():
    n = 70000                                            # > 65535 entries => ZIP64 EOCD record
    members = [synth.ZipMember("t%05d" % i, b"", method=0) for i in range(n)]
    z = synth.make_zip(members, threads=1)
    entries, _, _ = capi.zip_index(z)
    assert len(entries) == n and entries[-1]["name"] == b"t69999"


def test_zip_index_rejects_garbage():
    for blob in (b"", b"PK", b"\0" * 100, b"PK\x05\x06" + b"\0" * 18):   # EOCD at offset 0: i > 0 rule
        with pytest.raises(capi.B2IError):
            capi.zip_index(blob)


def test_local_values_win_and_warn():
    txt = synth.synth_text(2000, 9)
    z = bytearray(synth.make_zip([synth.ZipMember("a", txt)]))
    crc_pos = 14                                          # local header CRC field
    z[crc_pos] ^= 0xFF
    entries, _, _ = capi.zip_index(bytes(z))
    assert entries[0]["warn"] & 1 and entries[0]["crc32"] != zlib.crc32(txt)   # zip.c:1112-1119


def test_gzip_header_parsing_and_bgzf_chain():
    L = capi.lib()
    import ctypes as C
    data = synth.synth_text(1000, 1)
    for kw in ({}, {"name": b"file.txt"}, {"comment": b"hello"}, {"extra": b"AB\x01\x00x"}, {"hcrc": True},
               {"name": b"n", "comment": b"c", "extra": b"", "hcrc": True, "mtime": 77}):
        g = synth.gzip_member(data, **kw)
        m = capi.GzipMember()
        hl = L.b2i_gzip_peek_header(g, len(g), 0, C.byref(m))
        assert hl and zlib.decompress(g[hl:-8], -15) == data
        assert m.mtime == kw.get("mtime", 0) and m.deflate_len == 0      # no BSIZE: length unknown
    for bad in (b"\x1f\x8b\x07" + bytes(20), b"\x1f\x8b\x08\xe0" + bytes(20), b"\x1f\x8b\x08\x08" + b"a" * 20):
        assert L.b2i_gzip_peek_header(bad, len(bad), 0, None) == 0
    parts = [synth.synth_text(n, n) for n in (1, 100, 65280, 3000)]
    f = synth.make_bgzf(parts)
    members, end = capi.gzip_scan_bgzf(f + b"garbage")
    assert len(members) == 5 and end == len(f)                       # 4 data members + EOF marker
    for m, p in zip(members, parts + [b""]):
        raw = f[m["deflate_offset"]:m["deflate_offset"] + m["deflate_len"]]
        assert zlib.decompress(raw, -15) == p and m["crc32"] == zlib.crc32(p) and m["isize"] == len(p)


def test_plan_layout_and_partition():
    parts = synth.split_text(64 * 7001, 7001, 6)
    z = synth.make_zip([synth.ZipMember("p%02d" % i, p) for i, p in enumerate(parts)] +
                       [synth.ZipMember("s", b"y" * 999, method=0), synth.ZipMember("e", b"")])
    entries, _, _ = capi.zip_index(z)
    descs, out_bytes, which = reader.plan_zip(entries)
    assert len(descs) == 66 and all(d.out_off % 16 == 0 for d in descs)
    assert descs[64].flags & capi.F_NO_COPY and descs[64].out_cap == 0
    ends = sorted((int(d.out_off), int(d.out_off + d.out_cap)) for d in descs)
    assert all(a[1] <= b[0] for a, b in zip(ends, ends[1:])) and ends[-1][1] <= out_bytes
    for world in (1, 2, 3, 8):
        # the C partitioner of the library (b2i_partition_contiguous / b2i_partition_lpt)
        w = [int(d.in_len + max(d.out_cap, d.expect_out)) + 1 for d in descs]
        cuts = shard.partition_contiguous(descs, world)
        assert cuts[0][0] == 0 and cuts[-1][1] == len(w) and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        loads = [sum(w[lo:hi]) for lo, hi in cuts]
        assert max(loads) <= sum(w) / world + max(w)
        owner, load = capi.partition_lpt(descs, world)
        assert sorted(set(owner)) == list(range(min(world, len(w)))) and sum(load) == sum(w)
        assert max(load) <= sum(w) / world + max(w)


def test_zip_index_through_fetch_callback_equals_memory_walk():
    """File-backed sources: the index walk through bounded fetches (tail, directory, local
    headers) gives the same entries as the walk over a memory image, for plain, ZIP64,
    prefixed and damaged archives; the tail probe answers like the seekable bid."""
    from libarchive_b200 import synth
    txt = synth.synth_text(40000, 3)
    members = [synth.ZipMember("a/%03d.txt" % i, txt[i * 90:i * 90 + 300 + 7 * i]) for i in range(120)]
    cases = [synth.make_zip(members), synth.make_zip(members, zip64=True, framing="at_end"),
             synth.make_zip(members[:5], prefix=b"#!/bin/sh\n" * 40, comment=b"hello")]
    for z in cases:
        want = capi.zip_index(z)
        got = capi.zip_index_via_fetch(z, chunk=4096)
        strip = lambda es: [{k: v for k, v in e.items() if k != "reserved"} for e in es]
        assert strip(got[0]) == strip(want[0]) and got[1:3] == want[1:3]
        assert got[3] < len(z) + 16384 + 4096 * 2          # tail + directory + headers, not the file many times over
        tail = z[-16384:]
        assert capi.lib().b2i_zip_probe_tail(tail, len(tail), len(z)) == 1
    z = cases[0]
    with pytest.raises(capi.B2IError):                       # a zeroed end record
        capi.zip_index_via_fetch(z[:-30] + bytes(30))
    junk = synth.synth_random(50000, 9)
    assert capi.lib().b2i_zip_probe_tail(junk[-16384:], 16384, len(junk)) == 0
    with pytest.raises(capi.B2IError):
        capi.zip_index_via_fetch(junk)


def test_plugin_pinned_buffer_shelf(tmp_path):
    """The plugins' shelf of pinned output buffers (csrc/plugin/b200_ctx_pool.c): size
    classes, reuse, smallest-first eviction, bounded idle total - compiled with malloc in
    place of cudaHostAlloc (tests/csrc/buf_pool_test.c)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "buf_pool_test")
    subprocess.run(["gcc", "-O1", "-Wall", "-I" + os.path.join(root, "include"),
                    "-I" + os.path.join(root, "libarchive_b200", "csrc", "plugin"), "-o", exe,
                    os.path.join(root, "tests", "csrc", "buf_pool_test.c"),
                    os.path.join(root, "libarchive_b200", "csrc", "plugin", "b200_ctx_pool.c"), "-lpthread"],
                   check=True, timeout=120)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
