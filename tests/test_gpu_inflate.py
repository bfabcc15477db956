"""GPU parity: the CUDA inflate + fused CRC path (through the C ABI) against the
oracle on the same inputs, bit-exact (bytes, CRC, consumed/produced counts,
status, zlib error class)."""
import ctypes as C
import zlib

import numpy as np
import pytest

import oracle_binding as ob
from libarchive_b200 import capi, reader, synth
from libarchive_b200.capi import StreamDesc

pytestmark = pytest.mark.gpu


def run_streams(ctx, streams, caps=None, lead=0, gap=3):
    """Pack raw streams into one input buffer (arbitrary, unaligned offsets) and
    decode them in ONE device pass; returns (descs, gpu results, gpu out, oracle results, oracle out)."""
    blob = bytearray(b"\xAA" * lead)
    items, out = [], 0
    for k, s in enumerate(streams):
        d = StreamDesc()
        d.in_off, d.in_len, d.method = len(blob), len(s), 8
        d.out_off = out
        d.out_cap = caps[k] if caps else 1 << 18
        out = (out + d.out_cap + 15) & ~15
        blob += s + b"\x55" * gap
        items.append(d)
    descs = capi.make_descs(items)
    blob = bytes(blob)
    inbuf = C.create_string_buffer(blob, len(blob) + 32)
    outbuf = C.create_string_buffer(out + 32)
    res = ctx.decode_host(inbuf, len(blob), descs, outbuf, out)
    ores, oout = ob.decode_batch(blob, descs, out)
    return descs, res, outbuf.raw, ores, oout


def compare(names, descs, res, gout, ores, oout):
    for k, name in enumerate(names):
        r, o, d = res[k], ores[k], descs[k]
        assert r.status == o.status, (name, r.status, r.detail, o.status, o.detail)
        assert r.out_bytes == o.out_bytes, (name, r.out_bytes, o.out_bytes)
        a = gout[d.out_off:d.out_off + r.out_bytes]
        b = oout[d.out_off:d.out_off + o.out_bytes]
        assert a == b, (name, "bytes differ")
        if o.status == 0:
            assert r.in_bytes == o.in_bytes, (name, r.in_bytes, o.in_bytes)
            assert r.crc == o.crc == (zlib.crc32(b) & 0xFFFFFFFF), name
            assert r.flags == o.flags, (name, r.flags, o.flags)
        if o.status == -3:
            assert r.detail == o.detail, (name, r.detail, o.detail)


def test_zoo_matches_oracle(ctx):
    zoo = [(n, s) for n, s in synth.deflate_zoo()]
    names = [n for n, _ in zoo]
    for lead in (0, 7):
        out = run_streams(ctx, [s for _, s in zoo], lead=lead)
        compare(names, *out)


def test_random_dynamic_blocks(ctx):
    streams = [synth.random_dynamic_stream(seed, 2500) for seed in range(120)]
    out = run_streams(ctx, streams, lead=1, gap=1)
    compare(["rand%d" % i for i in range(len(streams))], *out)


def test_text_levels_and_strategies(ctx):
    txt = synth.synth_text(1 << 20, 7)
    rnd = synth.synth_random(300000, 8)
    streams, names = [], []
    for lvl in (1, 3, 6, 9):
        for n in (1, 100, 4096, 65536, 262144, 1 << 20):
            streams.append(synth.deflate_raw(txt[:n], lvl)); names.append("text l%d n%d" % (lvl, n))
    for strat, sn in ((zlib.Z_FIXED, "fixed"), (zlib.Z_HUFFMAN_ONLY, "huff"), (zlib.Z_RLE, "rle"), (zlib.Z_FILTERED, "filt")):
        streams.append(synth.deflate_raw(txt[:200000], 6, strat)); names.append(sn)
    streams.append(synth.deflate_raw(rnd, 6)); names.append("random->stored blocks")
    streams.append(synth.deflate_raw(rnd, 0)); names.append("level 0 stored")
    streams.append(synth.deflate_raw(bytes(1 << 20), 9)); names.append("zeros dist1")
    streams.append(synth.deflate_raw(b"ab" * 300000, 6)); names.append("abab dist2")
    streams.append(synth.deflate_raw(bytes(range(256)) * 3000, 6)); names.append("period 256")
    streams.append(synth.deflate_mixed([(txt[:50000], 6, zlib.Z_DEFAULT_STRATEGY), (rnd[:70000], 6, zlib.Z_DEFAULT_STRATEGY),
                                        (txt[50000:90000], 1, zlib.Z_FIXED)])); names.append("mixed blocks")
    caps = [(1 << 20) + 64] * len(streams)
    out = run_streams(ctx, streams, caps=caps, lead=3)
    compare(names, *out)


def test_truncated_and_trailing_junk(ctx):
    txt = synth.synth_text(100000, 9)
    full = synth.deflate_raw(txt, 6)
    streams, names = [], []
    for cut in (1, 2, 3, 10, 100, 1000, len(full) // 2, len(full) - 5, len(full) - 1):
        streams.append(full[:cut]); names.append("cut%d" % cut)
    streams.append(full + b"JUNKJUNK"); names.append("junk after")
    fx = synth.deflate_raw(txt[:3000], 6, zlib.Z_FIXED)
    for cut in range(1, 40):
        streams.append(fx[:cut]); names.append("fixed cut%d" % cut)
    out = run_streams(ctx, streams, caps=[1 << 17] * len(streams))
    compare(names, *out)
    descs, res = out[0], out[1]
    assert res[names.index("junk after")].flags & capi.R_IN_MISMATCH


def test_single_bit_corruptions(ctx):
    """Flip one bit at many positions of a dynamic stream: whatever zlib (oracle)
    says - error class, partial output, or a different valid decode - the GPU says too."""
    txt = synth.synth_text(20000, 11)
    s = bytearray(synth.deflate_raw(txt, 6))
    rng = np.random.default_rng(5)
    streams, names = [], []
    for pos in list(range(0, 200)) + [int(x) for x in rng.integers(0, len(s) * 8, 300)]:
        t = bytearray(s)
        t[pos >> 3] ^= 1 << (pos & 7)
        streams.append(bytes(t)); names.append("flip%d" % pos)
    out = run_streams(ctx, streams, caps=[1 << 16] * len(streams))
    compare(names, *out)


def test_output_capacity_overflow(ctx):
    """A stream that wants more room than its descriptor reserves stops with
    B2I_S_OUT_OVERFLOW and never writes past its capacity (device buffer is
    pre-filled with a sentinel through the explicit plan API)."""
    txt = synth.synth_text(50000, 12)
    s = synth.deflate_raw(txt, 6)
    caps = [1000, 49999, 50000]
    L = capi.lib()
    items, out = [], 0
    for k, cap in enumerate(caps):
        d = StreamDesc()
        d.in_off, d.in_len, d.method = 0, len(s), 8
        d.out_off, d.out_cap = out, cap
        out = (out + cap + 15 + 64) & ~15
        items.append(d)
    descs = capi.make_descs(items)
    d_in = L.b2i_device_alloc(ctx.h, len(s))
    d_out = L.b2i_device_alloc(ctx.h, out)
    sentinel = b"\xEE" * out
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_in, s, len(s)))
    ctx._check(L.b2i_memcpy_h2d(ctx.h, d_out, sentinel, out))
    plan = C.c_void_p()
    ctx._check(L.b2i_plan_create(ctx.h, descs, 3, C.byref(plan)))
    ctx._check(L.b2i_plan_launch(plan, d_in, len(s), d_out, out))
    res = (capi.StreamResult * 3)()
    ctx._check(L.b2i_plan_results(plan, res))
    host = C.create_string_buffer(out)
    ctx._check(L.b2i_memcpy_d2h(ctx.h, host, d_out, out))
    ctx.sync()
    L.b2i_plan_destroy(plan)
    L.b2i_device_free(ctx.h, d_in)
    L.b2i_device_free(ctx.h, d_out)
    assert [r.status for r in res] == [capi.S_OUT_OVERFLOW, capi.S_OUT_OVERFLOW, 0]
    g = host.raw
    for k, cap in enumerate(caps):
        o = descs[k].out_off
        assert g[o:o + res[k].out_bytes] == txt[:res[k].out_bytes]
        assert res[k].out_bytes <= cap
        assert g[o + cap:o + cap + 64] == b"\xEE" * 64, "wrote past capacity of stream %d" % k
    assert g[descs[2].out_off:descs[2].out_off + 50000] == txt


def test_many_streams_one_pass(ctx):
    """2048 x 16 KiB entries + a few large ones: persistent warps pull work largest-first."""
    parts = synth.split_text(2048 * 16384, 16384, 21)
    streams = [synth.deflate_raw(p, 6) for p in parts[:2048]]
    big = synth.synth_text(3 << 20, 22)
    streams += [synth.deflate_raw(big, 6), synth.deflate_raw(big[: 1 << 20], 1)]
    caps = [16384] * 2048 + [3 << 20, 1 << 20]
    descs, res, gout, ores, oout = run_streams(ctx, streams, caps=caps, gap=0)
    for k in range(len(streams)):
        assert res[k].status == 0 and res[k].flags == 0 or k >= 0 and res[k].status == 0
        assert res[k].crc == ores[k].crc and res[k].out_bytes == ores[k].out_bytes
    assert gout[:descs[-1].out_off + caps[-1]] == oout[:descs[-1].out_off + caps[-1]]


def test_fuzz_garbage_and_mutations(ctx):
    """Arbitrary bytes never hang or fault the kernel and always classify like zlib:
    pure garbage, garbage behind a valid dynamic header, and multi-bit mutations of
    long (lane-parallel sized) streams."""
    rng = np.random.default_rng(77)
    txt = synth.synth_text(200000, 13)
    base = [synth.deflate_raw(txt[:70000], 6), synth.deflate_raw(txt[:70000], 6, zlib.Z_FIXED),
            synth.deflate_raw(txt, 9), synth.random_dynamic_stream(5, 20000)]
    streams, names = [], []
    for k in range(300):
        n = int(rng.integers(1, 3000))
        streams.append(rng.integers(0, 256, n, dtype=np.uint8).tobytes()); names.append("garbage%d" % k)
    for k in range(400):
        s = bytearray(base[k % len(base)])
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, len(s) * 8))
            s[pos >> 3] ^= 1 << (pos & 7)
        if k % 5 == 0:
            s = s[:int(rng.integers(1, len(s)))]
        streams.append(bytes(s)); names.append("mut%d" % k)
    hdr = base[0][:120]
    for k in range(100):
        streams.append(hdr + rng.integers(0, 256, int(rng.integers(600, 6000)), dtype=np.uint8).tobytes())
        names.append("hdr+garbage%d" % k)
    out = run_streams(ctx, streams, caps=[1 << 18] * len(streams), lead=2, gap=5)
    compare(names, *out)


def test_pinned_host_buffers_pipelined_path(ctx, monkeypatch):
    """With pinned host buffers (b2i_host_alloc) b2i_decode_host pipelines slices over
    three streams (and, with B2I_MIRROR=1, lets the kernel store straight to host memory)
    - same bytes, same results either way."""
    L = capi.lib()
    parts = synth.split_text(600 * 65536, 65536, 91)
    big = synth.synth_text(3 << 20, 92)
    members = [synth.ZipMember("p%04d" % i, p) for i, p in enumerate(parts)]
    members += [synth.ZipMember("big", big), synth.ZipMember("rnd", synth.synth_random(200000, 3)),
                synth.ZipMember("fixed", big[:300000], strategy=zlib.Z_FIXED),
                synth.ZipMember("bad", parts[0], crc=1)]
    z = synth.make_zip(members)
    from libarchive_b200 import reader
    entries, _, _ = capi.zip_index(z)
    descs, out_bytes, which = reader.plan_zip(entries)
    n = len(descs)
    h_in = L.b2i_host_alloc(len(z) + 64)
    h_out = L.b2i_host_alloc(out_bytes + 64)
    C.memmove(h_in, z, len(z))
    C.memset(h_out, 0xEE, out_bytes + 64)
    res = (capi.StreamResult * n)()
    ores, oout = ob.decode_batch(z, descs, out_bytes)
    for mirror in (False, True, False):      # later calls reuse the arena and the streams
        if mirror:
            monkeypatch.setenv("B2I_MIRROR", "1")
        else:
            monkeypatch.delenv("B2I_MIRROR", raising=False)
        C.memset(h_out, 0xEE, out_bytes + 64)
        ctx._check(L.b2i_decode_host(ctx.h, h_in, len(z), descs, n, h_out, out_bytes, res))
        got = C.string_at(h_out, out_bytes + 64)
        for k in range(n):
            assert (res[k].status, res[k].crc, res[k].out_bytes, res[k].in_bytes, res[k].flags) == \
                   (ores[k].status, ores[k].crc, ores[k].out_bytes, ores[k].in_bytes, ores[k].flags), k
            o = descs[k].out_off
            assert got[o:o + res[k].out_bytes] == oout[o:o + res[k].out_bytes], k
        assert res[n - 1].flags & capi.R_CRC_MISMATCH
        assert got[out_bytes:out_bytes + 64] == b"\xEE" * 64
    L.b2i_host_free(h_in)
    L.b2i_host_free(h_out)


def test_block_type_changes_at_every_input_alignment(ctx):
    """Streams that switch between dynamic, stored (incl. empty stored blocks from
    Z_FULL_FLUSH) and fixed blocks, placed at all 128 phases of the 128-byte input
    ring segments: resuming after a stored block may step back into a ring half that
    has already been refilled (regression: found by the mixed-blocks config)."""
    txt = synth.synth_text(60000, 5)
    rnd = synth.synth_random(20000, 6)
    s1 = synth.deflate_mixed([(txt[:7000], 6, zlib.Z_DEFAULT_STRATEGY), (rnd[:9000], 6, zlib.Z_DEFAULT_STRATEGY),
                              (txt[7000:9000], 1, zlib.Z_FIXED), (rnd[:100], 6, zlib.Z_DEFAULT_STRATEGY),
                              (txt[9000:40000], 6, zlib.Z_DEFAULT_STRATEGY)])
    s2 = synth.deflate_raw(rnd, 6) + b""
    streams, names = [], []
    blob = bytearray()
    items, out = [], 0
    for phase in range(128):
        for s in (s1, s2):
            pad = (phase - len(blob)) % 128
            blob += b"\x00" * pad
            d = StreamDesc()
            d.in_off, d.in_len, d.method = len(blob), len(s), 8
            d.out_off, d.out_cap = out, 65536
            out += 65536
            blob += s
            items.append(d); names.append("phase%d" % phase)
    descs = capi.make_descs(items)
    blob = bytes(blob)
    inbuf = C.create_string_buffer(blob, len(blob) + 32)
    outbuf = C.create_string_buffer(out + 32)
    res = ctx.decode_host(inbuf, len(blob), descs, outbuf, out)
    ores, oout = ob.decode_batch(blob, descs, out)
    compare(names, descs, res, outbuf.raw, ores, oout)
    assert all(r.status == 0 for r in res)


def test_team_decoder_opt_in(ctx, monkeypatch):
    """B2I_TEAM_MIN_BYTES routes large streams to the experimental team kernel (one CTA
    of four warps per stream): identical results, including errors and capacity limits."""
    monkeypatch.setenv("B2I_TEAM_MIN_BYTES", "20000")
    txt = synth.synth_text(3 << 20, 17)
    rnd = synth.synth_random(300000, 18)
    full = synth.deflate_raw(txt[:600000], 6)
    flip = bytearray(full); flip[150000] ^= 0x20
    streams = [synth.deflate_raw(txt, 6), synth.deflate_raw(txt[:900000], 9), full[:100000], bytes(flip), full,
               synth.deflate_raw(txt[:700000], 1, zlib.Z_FIXED), synth.deflate_raw(rnd, 6),
               synth.deflate_raw(bytes(2 << 20), 6) + b"", synth.deflate_raw(b"ab" * 400000 + txt[:100000], 6),
               synth.deflate_mixed([(txt[:200000], 6, zlib.Z_DEFAULT_STRATEGY), (rnd[:90000], 6, zlib.Z_DEFAULT_STRATEGY),
                                    (txt[200000:500000], 1, zlib.Z_FIXED)]),
               synth.random_dynamic_stream(9, 40000), synth.deflate_raw(txt[:5000], 6)]
    names = ["text3M", "text900k-l9", "truncated", "bitflip", "cap-too-small", "fixed", "stored", "zeros", "abab",
             "mixed", "random-codes", "small"]
    caps = [(3 << 20) + 64] * len(streams)
    caps[4] = 250000
    out = run_streams(ctx, streams, caps=caps, lead=5, gap=3)
    compare(names, *out)


def test_two_jobs_in_flight(ctx):
    """b2i_submit / b2i_wait: two host-buffer decodes of one context overlap; results and
    bytes equal the one-call path and the oracle; a third submit is refused."""
    import ctypes as C
    L = capi.lib()
    batches = []
    for seed in (1, 2, 3):
        members = [synth.ZipMember("f%d_%d" % (seed, i), synth.synth_text(30000 + 1000 * i, 10 * seed + i)) for i in range(40)]
        members.append(synth.ZipMember("bad", synth.synth_text(5000, seed), crc=123))
        z = synth.make_zip(members)
        entries, _, _ = capi.zip_index(z)
        descs, out_bytes, _ = reader.plan_zip(entries, stored_no_copy=False)
        h_in = L.b2i_host_alloc(len(z) + 64)
        h_out = L.b2i_host_alloc(out_bytes + 64)
        C.memmove(h_in, z, len(z))
        batches.append((z, descs, out_bytes, h_in, h_out))
    jobs = [ctx.submit(b[3], len(b[0]), b[1], b[4], b[2]) for b in batches[:2]]
    # five jobs may be in flight on one context; a sixth is refused
    extra = [ctx.submit(batches[2][3], len(batches[2][0]), batches[2][1], batches[2][4], batches[2][2])
             for _ in range(3)]
    with pytest.raises(capi.B2IError):
        ctx.submit(batches[2][3], len(batches[2][0]), batches[2][1], batches[2][4], batches[2][2])
    for j in extra:
        ctx.wait(j)
    results = [ctx.wait(jobs[0])]
    jobs.append(ctx.submit(batches[2][3], len(batches[2][0]), batches[2][1], batches[2][4], batches[2][2]))
    results += [ctx.wait(jobs[1]), ctx.wait(jobs[2])]
    for (z, descs, out_bytes, h_in, h_out), res in zip(batches, results):
        ores, oout = ob.decode_batch(z, descs, out_bytes)
        got = C.string_at(h_out, out_bytes)
        for k in range(len(descs)):
            assert (res[k].status, res[k].crc, res[k].out_bytes, res[k].in_bytes, res[k].flags) == \
                   (ores[k].status, ores[k].crc, ores[k].out_bytes, ores[k].in_bytes, ores[k].flags), k
            o, nb = descs[k].out_off, res[k].out_bytes
            assert got[o:o + nb] == oout[o:o + nb], k
        L.b2i_host_free(h_in)
        L.b2i_host_free(h_out)


def test_nine_bit_root_kernel(ctx, monkeypatch):
    """The second build of the inflate kernel (b2i_inflate_kernel_r9: 9-bit lit/len root, 8 CTAs
    per SM; picked for batches with several waves of streams) forced onto the zoo, random codes,
    mutated streams and real text: identical results."""
    monkeypatch.setenv("B2I_KERNEL", "r9")
    zoo = [(n, s) for n, s in synth.deflate_zoo()]
    compare([n for n, _ in zoo], *run_streams(ctx, [s for _, s in zoo], lead=3))
    streams = [synth.random_dynamic_stream(seed, 2500) for seed in range(200, 280)]
    compare(["rand%d" % i for i in range(len(streams))], *run_streams(ctx, streams, lead=1, gap=1))
    rng = np.random.default_rng(5)
    txt = synth.synth_text(300000, 21)
    base = [synth.deflate_raw(txt[:70000], 6), synth.deflate_raw(txt, 9), synth.deflate_raw(txt[:70000], 6, zlib.Z_FIXED)]
    streams = list(base)
    for k in range(200):
        s = bytearray(base[k % len(base)])
        for _ in range(int(rng.integers(1, 6))):
            pos = int(rng.integers(0, len(s) * 8))
            s[pos >> 3] ^= 1 << (pos & 7)
        streams.append(bytes(s))
    compare(["m%d" % i for i in range(len(streams))], *run_streams(ctx, streams, caps=[1 << 19] * len(streams), lead=2, gap=5))


def test_many_waves_pick_the_second_build(ctx):
    """12 000 small streams (> 2 waves of 148 x 28 warps): the API picks the 8-CTA build by
    itself; results equal the oracle's."""
    txt = synth.synth_text(1 << 20, 4)
    streams = [synth.deflate_raw(txt[(i * 131) % 900000:][:2000 + (i % 7) * 300], 6) for i in range(12000)]
    out = run_streams(ctx, streams, caps=[8192] * len(streams), lead=0, gap=0)
    compare(["s%d" % i for i in range(len(streams))], *out)
