"""The oracle pinned: CPU restatement (oracle/*.c) against (1) Python's zlib - the
same zlib 1.3 the reference links - on generated and malformed streams, (2) the
reference test-suite's bit-at-a-time CRC, (3) the UNMODIFIED reference's results on
its own fixtures (tests/golden/ref_expected.json, produced by oracle/_ref)."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle_binding as ob
from libarchive_b200 import capi, reader, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def zl(stream):
    d = zlib.decompressobj(-15)
    try:
        out = d.decompress(stream)
    except zlib.error:
        return -3, None, None
    if d.eof:
        return 0, out, len(stream) - len(d.unused_data)
    return -5, out, len(stream)


def check(name, s, cap=1 << 21):
    r, out = ob.inflate(s, cap)
    st, zout, zin = zl(s)
    assert r.status == st, (name, r.status, r.detail, st)
    if st != -3:
        assert out == zout, name
    if st == 0:
        assert r.in_bytes == zin, name


def test_zoo_against_zlib():
    for name, s in synth.deflate_zoo():
        check(name, s)


def test_zoo_covers_every_zlib_message():
    seen = set()
    for name, s in synth.deflate_zoo():
        r, _ = ob.inflate(s, 1 << 20)
        if r.status == -3:
            seen.add(r.detail)
    assert seen == set(range(1, 12)), seen


def test_random_dynamic_blocks_against_zlib():
    for seed in range(60):
        check("rand%d" % seed, synth.random_dynamic_stream(seed, 2000))


def test_generated_streams_against_zlib():
    txt = synth.synth_text(600000, 1)
    for lvl in (0, 1, 6, 9):
        check("text l%d" % lvl, synth.deflate_raw(txt, lvl))
    check("fixed", synth.deflate_raw(txt[:100000], 6, zlib.Z_FIXED))
    check("random", synth.deflate_raw(synth.synth_random(200000), 6))
    check("zeros", synth.deflate_raw(bytes(500000), 9))
    full = synth.deflate_raw(txt[:100000], 6)
    for cut in (0, 1, 2, 5, 100, len(full) // 2, len(full) - 1):
        check("cut%d" % cut, full[:cut])
    check("junk", full + b"JUNK")
    rng = np.random.default_rng(1)
    for pos in rng.integers(0, len(full) * 8, 200):
        t = bytearray(full)
        t[int(pos) >> 3] ^= 1 << (int(pos) & 7)
        check("flip%d" % pos, bytes(t))


def test_crc32_semantics():
    L = ob.lib()
    assert L.orc_crc32(0, None, 0) == 0                 # archive_crc32.h:51-52
    assert ob.crc32(b"123456789") == 0xCBF43926
    rng = np.random.default_rng(2)
    for n in (0, 1, 7, 8, 9, 255, 256, 1000, 65537):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ob.crc32(d) == (zlib.crc32(d) & 0xFFFFFFFF) == L.orc_bitcrc32(0, d, n)
        assert ob.crc32(d[n // 2:], ob.crc32(d[:n // 2])) == ob.crc32(d)      # chaining
    for la, lb in [(0, 0), (0, 9), (9, 0), (1, 1), (100, 3), (4096, 65536), (70001, 12345)]:
        a = rng.integers(0, 256, la, dtype=np.uint8).tobytes()
        b = rng.integers(0, 256, lb, dtype=np.uint8).tobytes()
        assert L.orc_crc32_combine(ob.crc32(a), ob.crc32(b), lb) == ob.crc32(a + b)
        assert capi.lib().b2i_crc32_combine(ob.crc32(a), ob.crc32(b), lb) == ob.crc32(a + b)


def test_oracle_against_reference_fixtures():
    """Every deflate / stored entry of the reference's ZIP fixtures, located with the
    product's host index, decoded by the oracle: sizes and CRCs as the unmodified
    reference reported them."""
    with open(os.path.join(GOLD, "ref_expected.json")) as f:
        expected = json.load(f)
    checked = 0
    for name, exp in sorted(expected.items()):
        if exp["raw"]:
            continue
        with open(os.path.join(GOLD, "ref_fixtures", name), "rb") as f:
            blob = f.read()
        try:
            entries, _, _ = capi.zip_index(blob)
        except capi.B2IError:
            continue
        want = [r for r in exp["report"] if "i" in r and r.get("hdr", 0) != -30]
        descs, out_bytes, which = reader.plan_zip(entries, stored_no_copy=False)
        res, out = ob.decode_batch(blob, descs, out_bytes)
        for k, ei in enumerate(which):
            w = want[ei]
            if w["rd"] != 1 or (w["mode"] & 0o170000) != 0o100000:
                continue
            assert res[k].status == 0 and res[k].out_bytes == w["nbytes"], (name, ei)
            assert "%08x" % res[k].crc == w["crc"], (name, ei)
            checked += 1
    assert checked >= 40


def test_oracle_matches_reference_binary_when_present():
    """Where oracle/_ref exists (build container and GPU box), run the unmodified
    reference on a generated archive and compare every entry with the oracle."""
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    import tempfile
    parts = synth.split_text(24 * 30000, 30000, 5)
    members = [synth.ZipMember("t%02d" % i, p, level=1 + i % 9) for i, p in enumerate(parts)]
    members.append(synth.ZipMember("fixed", parts[0], strategy=zlib.Z_FIXED))
    members.append(synth.ZipMember("stored", synth.synth_random(40000, 1), method=0))
    members.append(synth.ZipMember("rnd-deflate", synth.synth_random(40000, 2)))
    z = synth.make_zip(members, framing="at_end")
    with tempfile.NamedTemporaryFile(suffix=".zip", delete=False) as f:
        f.write(z)
    try:
        lines, data = ob.ref_list(f.name)
    finally:
        os.unlink(f.name)
    entries, _, _ = capi.zip_index(z)
    descs, out_bytes, which = reader.plan_zip(entries, stored_no_copy=False)
    res, out = ob.decode_batch(z, descs, out_bytes)
    rows = [l for l in lines if "i" in l]
    assert len(rows) == len(members)
    for k, ei in enumerate(which):
        assert rows[ei]["rd"] == 1 and rows[ei]["nbytes"] == res[k].out_bytes
        assert rows[ei]["crc"] == "%08x" % res[k].crc
    assert hashlib.sha256(data).hexdigest() == hashlib.sha256(b"".join(m.data for m in members)).hexdigest()
