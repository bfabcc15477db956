"""not gpu: the streaming engine's HOST logic (csrc/b2i_pipe.cpp: window cutting, the pinned
ring, worker threads, release / skip rules, several contexts, error propagation) through its
C ABI, as compiled into tests/refsuite/_out/libarchive_hostlogic.so - device calls answered by
the oracle shim (tests/emul/b2i_shim.c, test infrastructure only).  The same scenarios run on
the GPU in test_gpu_pipe.py."""
import ctypes as C
import os

import pytest

from libarchive_b200 import capi
import test_gpu_pipe as scenarios

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "tests", "refsuite", "_out", "libarchive_hostlogic.so")

pytestmark = pytest.mark.timeout(300, method="thread")


@pytest.fixture(scope="module")
def H():
    if not os.path.exists(HOSTLIB):
        pytest.skip("tests/refsuite/_out/libarchive_hostlogic.so not built (needs /root/reference)")
    L = C.CDLL(HOSTLIB)
    vp, sz = C.c_void_p, C.c_size_t
    L.b2i_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.b2i_ctx_destroy.argtypes = [vp]
    L.b2i_ctx_destroy.restype = None
    L.b2i_pipe_open.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_uint64, capi.FILL_FN, vp, C.POINTER(capi.StreamDesc), sz,
                                C.POINTER(capi.PipeOpts), C.POINTER(vp)]
    L.b2i_pipe_get.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(vp), C.POINTER(capi.StreamResult)]
    L.b2i_pipe_release.argtypes = [vp, sz]
    L.b2i_pipe_release.restype = None
    L.b2i_pipe_window_count.argtypes = [vp]
    L.b2i_pipe_window_count.restype = sz
    L.b2i_pipe_error.argtypes = [vp]
    L.b2i_pipe_error.restype = C.c_char_p
    L.b2i_pipe_close.argtypes = [vp]
    L.b2i_pipe_close.restype = None
    return L


class HostCtx:
    def __init__(self, L):
        self.L, self.h = L, C.c_void_p()
        assert L.b2i_ctx_create(0, None, C.byref(self.h)) == 0

    def close(self):
        self.L.b2i_ctx_destroy(self.h)


class HostPipe(capi.Pipe):
    """capi.Pipe bound to the host-logic library instead of libb200inflate.so"""

    def __init__(self, L, ctxs, descs, mem=None, mem_size=0, fill=None, window_out=0, first_window_out=0, depth=0):
        self.L, self.n, self._keep = L, len(descs), (mem, descs)
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        opts = capi.PipeOpts(window_out, first_window_out, depth, 0)
        self._cb = capi.FILL_FN(fill) if fill is not None else C.cast(None, capi.FILL_FN)
        h = C.c_void_p()
        rc = L.b2i_pipe_open(arr, len(ctxs), C.cast(mem, C.c_void_p) if mem is not None else None, mem_size, self._cb,
                             None, descs, self.n, C.byref(opts), C.byref(h))
        if rc != capi.OK:
            raise capi.B2IError(f"b2i_pipe_open: {rc}")
        self.h = h


def test_memory_source_windows_and_two_contexts(H):
    z, members, descs = scenarios._archive()
    buf = C.create_string_buffer(z, len(z) + 64)
    c1, c2 = HostCtx(H), HostCtx(H)
    p = HostPipe(H, [c1], descs, mem=buf, mem_size=len(z), window_out=1 << 20, first_window_out=1 << 18)
    assert p.windows > 4
    scenarios._check_all(p, members, descs)
    p.close()
    p = HostPipe(H, [c1, c2], descs, mem=buf, mem_size=len(z), window_out=1 << 19, first_window_out=1 << 18, depth=2)
    scenarios._check_all(p, members, descs)
    p.close()
    c1.close(); c2.close()


def test_callback_source_skip_and_failing_source(H):
    z, members, descs = scenarios._archive(n=120, seed=9)
    calls = []

    def fill(user, off, length, dst):
        calls.append((off, length))
        C.memmove(dst, z[off:off + length], length)
        return 0

    c = HostCtx(H)
    p = HostPipe(H, [c], descs, fill=fill, window_out=1 << 20, first_window_out=1 << 19, depth=2)
    scenarios._check_all(p, members, descs, order=[0, 1, 2, 50, 51, 119])    # jumps drop unstarted windows
    p.close()
    assert calls and sum(l for _, l in calls) < len(z) + (1 << 20)
    p = HostPipe(H, [c], descs, fill=lambda u, o, l, d: -5, window_out=1 << 20)
    with pytest.raises(capi.B2IError):
        p.get(0)
    p.close()
    # going back to a window that was given up is refused, not served from a recycled slot
    p = HostPipe(H, [c], descs, fill=fill, window_out=1 << 19, first_window_out=1 << 18, depth=2)
    p.get(len(descs) - 1)
    with pytest.raises(capi.B2IError):
        p.get(0)
    p.close()
    c.close()


def test_default_window_sizes(H):
    """The default policy (b2i_pipe_open): a quarter of the batch within 16..256 MiB, 32 x the
    largest stream up to 512 MiB for memory sources, 64 MiB at most for callback sources, the
    first window at most 16 MiB.  Descriptors only - every stream points at the same few
    (invalid) input bytes, nothing is consumed."""
    mem = C.create_string_buffer(4096)
    c = HostCtx(H)

    def descs_of(sizes):
        arr = (capi.StreamDesc * len(sizes))()
        off = 0
        for i, s in enumerate(sizes):
            arr[i].in_off, arr[i].in_len, arr[i].out_off, arr[i].out_cap, arr[i].expect_out = 0, 64, off, s, s
            arr[i].method = 8
            off += (s + 15) & ~15
        return arr

    def windows(sizes, **kw):
        p = HostPipe(H, [c], descs_of(sizes), **kw)
        n = p.windows
        p.close()
        return n

    MiB = 1 << 20
    # config-1 shape: 4096 x 64 KiB = 256 MiB -> 64 MiB windows, the first one 16 MiB
    assert windows([65536] * 4096, mem=mem, mem_size=4096) == 1 + -(-(256 - 16) // 64)
    # 2 GiB with 16 MiB entries among small ones: 512 MiB windows (32 x the largest)
    assert windows([16 * MiB] * 8 + [MiB] * 1920, mem=mem, mem_size=4096) == 1 + -(-(2048 - 16) // 512)
    # the same through a callback source: 256 MiB windows (16 x the largest); 64 MiB when the
    # entries are small
    assert windows([16 * MiB] * 8 + [MiB] * 1920, fill=lambda u, o, l, d: 0) == 1 + -(-(2048 - 16) // 256)
    assert windows([MiB] * 2048, fill=lambda u, o, l, d: 0) == 1 + -(-(2048 - 16) // 64)
    # small archive: one 16 MiB window minimum -> first 4 MiB
    assert windows([65536] * 128, mem=mem, mem_size=4096) == 2
    c.close()
