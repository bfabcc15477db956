"""The warp-level device code (table construction, batch resolution, ring logic,
CRC merge tree) compiled for the host under tests/emul (32 pthreads = 1 warp) and
checked against the oracle - the not-gpu suite's coverage of the kernel logic.
The TMA/mbarrier PTX itself is only exercised by the -m gpu tests."""
import ctypes as C
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle_binding as ob
import test_gpu_reader as gr
from emul_ctx import EmulContext, emu
from libarchive_b200 import capi, reader, synth
from libarchive_b200.capi import StreamDesc, StreamResult


def emul_inflate(s, cap, lead=0):
    buf = bytes(lead) + s
    inb = C.create_string_buffer(buf + bytes(48), len(buf) + 48)
    out = C.create_string_buffer(cap + 64)
    d = StreamDesc()
    d.in_off, d.in_len, d.out_cap, d.method = lead, len(s), cap, 8
    r = StreamResult()
    emu().emul_inflate(inb, len(buf), out, C.byref(d), C.byref(r))
    return r, out.raw[:r.out_bytes]


def check(name, s, cap=1 << 18, leads=(0, 5)):
    o, od = ob.inflate(s, cap)
    for lead in leads:
        r, ed = emul_inflate(s, cap, lead)
        assert r.status == o.status, (name, lead, r.status, r.detail, o.status, o.detail)
        assert ed == od, (name, lead)
        if o.status == 0:
            assert r.in_bytes == o.in_bytes and r.crc == (zlib.crc32(od) & 0xFFFFFFFF), name
        if o.status == -3:
            assert r.detail == o.detail, name


def test_zoo():
    for name, s in synth.deflate_zoo():
        check(name, s)


def test_random_dynamic_blocks():
    for seed in range(12):
        check("rand%d" % seed, synth.random_dynamic_stream(100 + seed, 1200), leads=(seed % 16,))


def test_text_fixed_stored_and_errors():
    txt = synth.synth_text(70000, 2)
    full = synth.deflate_raw(txt[:65536], 6)
    check("text", full, leads=(3,))
    check("fixed", synth.deflate_raw(txt[:9000], 6, zlib.Z_FIXED), leads=(1,))
    check("stored", synth.deflate_raw(synth.synth_random(40000), 6), leads=(15,))
    check("rle", synth.deflate_raw(bytes(70000), 6), leads=(0,))
    check("cut", full[:4000], leads=(2,))
    check("junk", synth.deflate_raw(txt[:800], 6) + b"JUNKJUNK", leads=(0,))
    r, out = emul_inflate(synth.deflate_raw(txt[:5000], 6), 1000)
    assert r.status == capi.S_OUT_OVERFLOW and r.out_bytes <= 1000 and out == txt[:r.out_bytes]


def test_crc_lengths_alignments():
    rng = np.random.default_rng(3)
    for n in [0, 1, 15, 16, 17, 31, 32, 100, 511, 512, 513, 1000, 4096, 65536, 65537, 100001]:
        for off in (0, 3):
            data = rng.integers(0, 256, n + off + 64, dtype=np.uint8).tobytes()
            buf = C.create_string_buffer(data, len(data))
            assert emu().emul_crc32(0, buf, off, n) == (zlib.crc32(data[off:off + n]) & 0xFFFFFFFF), (n, off)
            assert emu().emul_crc32(0xABCD, buf, off, n) == (zlib.crc32(data[off:off + n], 0xABCD) & 0xFFFFFFFF)


@pytest.mark.parametrize("name", gr.ZIP_NAMES)
def test_reader_on_reference_zip_fixtures(name):
    gr.test_reference_zip_fixture(EmulContext(), name)


@pytest.mark.parametrize("name", gr.GZ_NAMES)
def test_reader_on_reference_gzip_fixtures(name):
    gr.test_reference_gzip_fixture(EmulContext(), name)


def test_reader_error_messages():
    gr.test_entry_errors_match_reference_messages(EmulContext())
    gr.test_ignorecrc32_option(EmulContext())


def test_stored_block_resume_at_every_ring_phase():
    """Regression (found on the GPU by the mixed-blocks config): after a stored block the
    bit reader resumes at a byte position up to 8 bytes behind what it had loaded; when
    that steps back across a 128-byte ring segment the half has already been refilled."""
    txt = synth.synth_text(1500, 5)
    rnd = synth.synth_random(300, 6)
    s = synth.deflate_mixed([(txt[:600], 6, zlib.Z_DEFAULT_STRATEGY), (rnd[:250], 6, zlib.Z_DEFAULT_STRATEGY),
                             (txt[600:800], 1, zlib.Z_FIXED), (rnd[:60], 6, zlib.Z_DEFAULT_STRATEGY),
                             (txt[800:], 6, zlib.Z_DEFAULT_STRATEGY)])
    check("mixed", s, cap=1 << 14, leads=range(0, 128))


def test_team_decoder_cases():
    """inflate_team.cuh under the emulator (8 warps = 256 pthreads): shared tables,
    256-segment rounds, tokens -> bytes through the shared-memory chunk buffer (expand,
    pointer sweep, flush + history ring), stored blocks copied by the team, team CRC,
    partial regions (one segment larger than the chunk), capacity and distance errors cut
    at the exact symbol."""
    txt = synth.synth_text(120000, 23)

    def team(s, cap, lead):
        buf = bytes(lead) + s
        inb = C.create_string_buffer(buf + bytes(48), len(buf) + 48)
        out = C.create_string_buffer(cap + 64)
        d = StreamDesc()
        d.in_off, d.in_len, d.out_cap, d.method = lead, len(s), cap, 8
        r = StreamResult()
        emu().emul_inflate_team(inb, len(buf), out, C.byref(d), C.byref(r))
        return r, out.raw[:r.out_bytes]

    full = synth.deflate_raw(txt, 6)
    dic = synth.synth_text(30000, 77)
    far_body = synth.synth_text(20000, 78) + dic[25000:] + synth.synth_text(9000, 79)
    co = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_DEFAULT_STRATEGY, dic)
    far = synth.deflate_raw(txt[:2000], 6)      # placeholder replaced below
    c1 = zlib.compressobj(6, zlib.DEFLATED, -15)
    far = c1.compress(txt[:2000]) + c1.flush(zlib.Z_FULL_FLUSH) + co.compress(far_body) + co.flush()
    mixed = synth.deflate_mixed([(txt[:50000], 6, zlib.Z_DEFAULT_STRATEGY),
                                 (synth.synth_random(70000, 3), 6, zlib.Z_DEFAULT_STRATEGY),
                                 (txt[50000:100000], 1, zlib.Z_FIXED)])
    cases = [("text", full, 1 << 19, 9), ("small", synth.deflate_raw(txt[:14000], 6), 1 << 15, 3),
             ("truncated", full[:25000], 1 << 19, 0), ("cap", full, 63457, 5),
             ("fixed", synth.deflate_raw(txt[:70000], 1, zlib.Z_FIXED), 1 << 18, 1),
             ("stored-blocks", synth.deflate_raw(synth.synth_random(200000, 5), 6), 1 << 18, 7),
             # n = 8 * 16 * k + 3: the eight CRC slices must still cover the last bytes
             ("crc-slices", synth.deflate_raw(synth.synth_random(8 * 16 * 600 + 3, 6), 6), 1 << 17, 6),
             ("mixed", mixed, 1 << 19, 11),
             ("rle", synth.deflate_raw((b"ab" * 50000 + txt[:5000]) * 2, 6), 1 << 20, 2),
             ("zeros", synth.deflate_raw(bytes(5 << 20), 6), 6 << 20, 13),
             ("zeros-cap", synth.deflate_raw(bytes(5 << 20), 6), (5 << 20) - 100001, 13),
             ("far-mid-stream", far, 1 << 17, 4)]
    for name, s, cap, lead in cases:
        o, od = ob.inflate(s, cap)
        r, ed = team(s, cap, lead)
        assert r.status == o.status and r.out_bytes == o.out_bytes and ed == od, (name, r.status, o.status, r.out_bytes, o.out_bytes)
        if o.status == 0:
            assert r.in_bytes == o.in_bytes and r.crc == (zlib.crc32(od) & 0xFFFFFFFF), name
        if name == "far-mid-stream":
            assert o.status == -3 and 2000 < o.out_bytes < 30000 and r.detail == 11


def test_nine_bit_root_build(monkeypatch):
    """The second build of the inflate kernel (B2I_R9: 9-bit lit/len root, 856-entry table):
    same streams, same answers - including codes with 10..15-bit symbols that now live in
    secondary tables."""
    import emul_ctx
    monkeypatch.setattr(emul_ctx, "VARIANT", "r9")
    for name, s in synth.deflate_zoo():
        check(name, s, leads=(0,))
    for seed in range(12):
        check("rand%d" % seed, synth.random_dynamic_stream(300 + seed, 1200), leads=(seed % 16,))
    txt = synth.synth_text(70000, 5)
    check("text", synth.deflate_raw(txt[:65536], 9), leads=(7,))
    check("fixed", synth.deflate_raw(txt[:9000], 6, zlib.Z_FIXED), leads=(1,))
    check("cut", synth.deflate_raw(txt[:30000], 6)[:4000], leads=(2,))
