"""-m gpu: the streaming engine (b2i_pipe_*) and the multi-context call (b2i_decode_host_multi)
through the C ABI: windows through the pinned ring, memory and callback sources, skipping,
oversized streams, stored entries served from the staged input, several contexts."""
import ctypes as C
import zlib

import pytest

from libarchive_b200 import capi, reader, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(150, method="thread")]


def _archive(n=150, each=40000, seed=7, stored_every=7):
    parts = synth.split_text(n * each, each, seed)
    members = []
    for i, p in enumerate(parts):
        if stored_every and i % stored_every == 3:
            members.append(synth.ZipMember("s%04d" % i, synth.synth_random(each // 2 + i, i), method=0))
        else:
            members.append(synth.ZipMember("e%04d" % i, p[:each - 13 * (i % 5)]))
    z = synth.make_zip(members, threads=4)
    entries, _, _ = capi.zip_index(z)
    descs, out_bytes, which = reader.plan_zip(entries)          # stored entries: NO_COPY
    return z, members, descs


def _check_all(pipe, members, descs, order=None):
    n = len(descs)
    for k in (order or range(n)):
        out, inp, res = pipe.get(k)
        assert res.status == 0 and res.flags == 0, (k, res.status, res.flags)
        m = members[k]
        if m.method == 0:
            assert out is None and C.string_at(inp, len(m.data)) == m.data
            assert res.crc == (zlib.crc32(m.data) & 0xFFFFFFFF)
        else:
            assert res.out_bytes == len(m.data) and C.string_at(out, len(m.data)) == m.data
        pipe.release(k)


def test_pipe_memory_source_windows(ctx):
    z, members, descs = _archive()
    buf = C.create_string_buffer(z, len(z) + 64)                 # pageable: staged by the copy threads
    p = capi.Pipe([ctx], descs, mem=buf, mem_size=len(z), window_out=1 << 20, first_window_out=1 << 18)
    assert p.windows > 4
    _check_all(p, members, descs)
    p.close()


def test_pipe_callback_source_and_skip(ctx):
    z, members, descs = _archive(n=120, seed=9)
    calls = []

    def fill(user, off, length, dst):
        calls.append((off, length))
        C.memmove(dst, z[off:off + length], length)
        return 0

    p = capi.Pipe([ctx], descs, fill=fill, window_out=1 << 20, first_window_out=1 << 19, depth=2)
    _check_all(p, members, descs, order=[0, 1, 2, 50, 51, 119])   # jumps drop unstarted windows
    p.close()
    assert calls and sum(l for _, l in calls) < len(z) + (1 << 20)
    # a failing source surfaces as an error from get()
    p = capi.Pipe([ctx], descs, fill=lambda u, o, l, d: -5, window_out=1 << 20)
    with pytest.raises(capi.B2IError):
        p.get(0)
    p.close()


def test_pipe_oversized_stream_and_pinned_source(ctx):
    big = synth.synth_text(3 << 20, 3)
    members = [synth.ZipMember("a", big[:1000]), synth.ZipMember("big", big), synth.ZipMember("b", big[5000:9000])]
    z = synth.make_zip(members, threads=1)
    entries, _, _ = capi.zip_index(z)
    descs, _, _ = reader.plan_zip(entries)
    L = capi.lib()
    h = L.b2i_host_alloc(len(z) + 64)                            # pinned: used in place, no staging copy
    C.memmove(h, z, len(z))
    p = capi.Pipe([ctx], descs, mem=h, mem_size=len(z), window_out=1 << 20)
    assert p.windows == 3
    _check_all(p, members, descs)
    p.close()
    L.b2i_host_free(h)


def test_pipe_two_contexts_round_robin(ctx):
    """Windows of one archive go round-robin over the contexts (one per GPU when the box has
    several; two contexts on one device otherwise): same bytes, in order."""
    L = capi.lib()
    ctx2 = capi.Context(1 if L.b2i_device_count() > 1 else 0)
    z, members, descs = _archive(n=260, seed=13, stored_every=9)
    buf = C.create_string_buffer(z, len(z) + 64)
    p = capi.Pipe([ctx, ctx2], descs, mem=buf, mem_size=len(z), window_out=1 << 20, first_window_out=1 << 19)
    assert p.windows > 6
    _check_all(p, members, descs)
    p.close()
    ctx2.close()


def test_decode_host_multi_two_contexts(ctx):
    """Two contexts (here on the same device) share one batch: contiguous partition for a
    uniform batch, LPT when one stream dominates; results and bytes equal zlib's."""
    ctx2 = capi.Context(0)
    L = capi.lib()
    for big in (0, 6 << 20):
        parts = synth.split_text(60 * 30000, 30000, 21)
        members = [synth.ZipMember("e%03d" % i, p) for i, p in enumerate(parts)]
        if big:
            members.insert(17, synth.ZipMember("huge", synth.synth_text(big, 5)))
        z = synth.make_zip(members, threads=4)
        entries, _, _ = capi.zip_index(z)
        descs, out_bytes, _ = reader.plan_zip(entries, stored_no_copy=False)
        n = len(descs)
        inbuf = C.create_string_buffer(z, len(z) + 64)
        outbuf = C.create_string_buffer(b"\x5a" * (out_bytes + 64), out_bytes + 64)
        res = (capi.StreamResult * n)()
        arr = (C.c_void_p * 2)(ctx.h, ctx2.h)
        rc = L.b2i_decode_host_multi(arr, 2, inbuf, len(z), descs, n, outbuf, out_bytes, res)
        assert rc == 0
        raw = outbuf.raw
        for k, m in enumerate(members):
            assert res[k].status == 0 and res[k].flags == 0
            assert raw[descs[k].out_off:descs[k].out_off + len(m.data)] == m.data
    ctx2.close()
