"""ctypes binding of include/b200inflate.h (libb200inflate.so, built in-tree).

This is the same C ABI the libarchive plugins bind; Python is only the test
and benchmark harness.  Loading fails loudly when the shared library is
missing: there is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2I_LIB") or os.path.join(_HERE, "libb200inflate.so")

OK, E_INVAL, E_CUDA, E_NOMEM, E_NODEVICE, E_FORMAT = 0, -1, -2, -3, -4, -5
S_OK, S_DATA_ERROR, S_BUF_ERROR, S_OUT_OVERFLOW, S_UNSUPPORTED = 0, -3, -5, -100, -101
F_NO_COPY, F_NO_CRC = 1, 2
R_CRC_MISMATCH, R_IN_MISMATCH, R_OUT_MISMATCH = 1, 2, 4
METHOD_STORED, METHOD_DEFLATE = 0, 8


class StreamDesc(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("in_len", C.c_uint64), ("out_off", C.c_uint64),
                ("out_cap", C.c_uint64), ("expect_out", C.c_uint64), ("expect_crc", C.c_uint32),
                ("method", C.c_uint8), ("flags", C.c_uint8), ("reserved", C.c_uint16)]


class StreamResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("crc", C.c_uint32), ("out_bytes", C.c_uint64),
                ("in_bytes", C.c_uint64), ("detail", C.c_uint32), ("flags", C.c_uint32)]


class ZipEntry(C.Structure):
    _fields_ = [("local_header_offset", C.c_uint64), ("data_offset", C.c_uint64),
                ("compressed_size", C.c_uint64), ("uncompressed_size", C.c_uint64),
                ("crc32", C.c_uint32), ("name_offset", C.c_uint32), ("name_len", C.c_uint16),
                ("zip_flags", C.c_uint16), ("method", C.c_uint16), ("version", C.c_uint8),
                ("system", C.c_uint8), ("mode", C.c_uint32), ("warn", C.c_uint32),
                ("mtime", C.c_int64), ("atime", C.c_int64), ("ctime", C.c_int64),
                ("uid", C.c_uint32), ("gid", C.c_uint32), ("local_extra_offset", C.c_uint64),
                ("local_extra_len", C.c_uint16), ("reserved", C.c_uint16 * 3)]


class ZipIndex(C.Structure):
    _fields_ = [("n", C.c_size_t), ("entries", C.POINTER(ZipEntry)), ("names", C.c_void_p),
                ("names_len", C.c_size_t), ("correction", C.c_int64),
                ("has_encrypted_entries", C.c_int)]


class GzipMember(C.Structure):
    _fields_ = [("header_offset", C.c_uint64), ("header_len", C.c_uint32),
                ("deflate_offset", C.c_uint64), ("deflate_len", C.c_uint64),
                ("crc32", C.c_uint32), ("isize", C.c_uint32), ("mtime", C.c_uint32),
                ("name_offset", C.c_uint32)]


class PipeOpts(C.Structure):
    _fields_ = [("window_out_bytes", C.c_size_t), ("first_window_out_bytes", C.c_size_t),
                ("windows_per_device", C.c_int), ("copy_threads", C.c_int)]


FILL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p)
FETCH_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t)

EXPORTS = [
    "b2i_ctx_create", "b2i_ctx_destroy", "b2i_last_error", "b2i_abi_version", "b2i_device_count",
    "b2i_ctx_sync", "b2i_ctx_launch_count", "b2i_host_alloc", "b2i_host_free", "b2i_device_alloc",
    "b2i_device_free", "b2i_memcpy_h2d", "b2i_memcpy_d2h", "b2i_plan_create", "b2i_plan_launch",
    "b2i_plan_results", "b2i_plan_destroy", "b2i_decode_host", "b2i_submit", "b2i_wait", "b2i_crc32", "b2i_crc32_device",
    "b2i_crc32_combine", "b2i_zip_index_build", "b2i_zip_index_free", "b2i_gzip_peek_header",
    "b2i_gzip_scan_bgzf", "b2i_free", "b2i_partition_contiguous", "b2i_partition_lpt",
    "b2i_decode_host_multi", "b2i_pipe_open", "b2i_pipe_get", "b2i_pipe_release", "b2i_pipe_window_count",
    "b2i_pipe_error", "b2i_pipe_close", "b2i_zip_index_build_cb", "b2i_zip_probe_tail", "b2i_ctx_device", "b2i_job_wait_input",
    "b2i_job_staged_input",
]

_lib = None


def lib():
    """Load libb200inflate.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C libarchive_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t
    L.b2i_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.b2i_ctx_destroy.argtypes = [vp]
    L.b2i_ctx_destroy.restype = None
    L.b2i_last_error.argtypes = [vp]
    L.b2i_last_error.restype = C.c_char_p
    L.b2i_ctx_sync.argtypes = [vp]
    L.b2i_ctx_device.argtypes = [vp]
    L.b2i_ctx_launch_count.argtypes = [vp]
    L.b2i_ctx_launch_count.restype = u64
    L.b2i_host_alloc.argtypes = [sz]
    L.b2i_host_alloc.restype = vp
    L.b2i_host_free.argtypes = [vp]
    L.b2i_host_free.restype = None
    L.b2i_device_alloc.argtypes = [vp, sz]
    L.b2i_device_alloc.restype = vp
    L.b2i_device_free.argtypes = [vp, vp]
    L.b2i_device_free.restype = None
    L.b2i_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.b2i_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.b2i_plan_create.argtypes = [vp, C.POINTER(StreamDesc), sz, C.POINTER(vp)]
    L.b2i_plan_launch.argtypes = [vp, vp, sz, vp, sz]
    L.b2i_plan_results.argtypes = [vp, C.POINTER(StreamResult)]
    L.b2i_plan_destroy.argtypes = [vp]
    L.b2i_plan_destroy.restype = None
    L.b2i_decode_host.argtypes = [vp, vp, sz, C.POINTER(StreamDesc), sz, vp, sz, C.POINTER(StreamResult)]
    L.b2i_submit.argtypes = [vp, vp, sz, C.POINTER(StreamDesc), sz, vp, sz, C.POINTER(vp)]
    L.b2i_wait.argtypes = [vp, C.POINTER(StreamResult)]
    L.b2i_job_wait_input.argtypes = [vp]
    L.b2i_job_staged_input.argtypes = [vp]
    L.b2i_job_staged_input.restype = vp
    L.b2i_crc32.argtypes = [vp, u32, vp, sz, C.POINTER(u32)]
    L.b2i_crc32_device.argtypes = [vp, u32, vp, sz, C.POINTER(u32)]
    L.b2i_crc32_combine.argtypes = [u32, u32, u64]
    L.b2i_crc32_combine.restype = u32
    L.b2i_zip_index_build.argtypes = [vp, sz, C.POINTER(ZipIndex), C.c_char_p]
    L.b2i_zip_index_free.argtypes = [C.POINTER(ZipIndex)]
    L.b2i_zip_index_free.restype = None
    L.b2i_gzip_peek_header.argtypes = [vp, sz, sz, C.POINTER(GzipMember)]
    L.b2i_gzip_peek_header.restype = sz
    L.b2i_gzip_scan_bgzf.argtypes = [vp, sz, sz, C.POINTER(C.POINTER(GzipMember)), C.POINTER(sz),
                                     C.POINTER(sz)]
    L.b2i_free.argtypes = [vp]
    L.b2i_free.restype = None
    L.b2i_partition_contiguous.argtypes = [C.POINTER(StreamDesc), sz, C.c_int, C.POINTER(sz)]
    L.b2i_partition_lpt.argtypes = [C.POINTER(StreamDesc), sz, C.c_int, C.POINTER(u32), C.POINTER(u64)]
    L.b2i_zip_index_build_cb.argtypes = [FETCH_FN, vp, u64, C.POINTER(ZipIndex), C.c_char_p]
    L.b2i_zip_probe_tail.argtypes = [vp, sz, u64]
    L.b2i_decode_host_multi.argtypes = [C.POINTER(vp), C.c_int, vp, sz, C.POINTER(StreamDesc), sz, vp, sz,
                                        C.POINTER(StreamResult)]
    L.b2i_pipe_open.argtypes = [C.POINTER(vp), C.c_int, vp, u64, FILL_FN, vp, C.POINTER(StreamDesc), sz,
                                C.POINTER(PipeOpts), C.POINTER(vp)]
    L.b2i_pipe_get.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(vp), C.POINTER(StreamResult)]
    L.b2i_pipe_release.argtypes = [vp, sz]
    L.b2i_pipe_release.restype = None
    L.b2i_pipe_window_count.argtypes = [vp]
    L.b2i_pipe_window_count.restype = sz
    L.b2i_pipe_error.argtypes = [vp]
    L.b2i_pipe_error.restype = C.c_char_p
    L.b2i_pipe_close.argtypes = [vp]
    L.b2i_pipe_close.restype = None
    _lib = L
    return L


class B2IError(RuntimeError):
    pass


class Context:
    """One b2i_ctx (one GPU, one stream)."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.b2i_ctx_create(device, C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(h))
        if rc != OK:
            raise B2IError(f"b2i_ctx_create(device={device}) failed: {rc} "
                           "(no usable sm_100 GPU; this library has no CPU path)")
        self.h = h

    def close(self):
        if self.h:
            self.L.b2i_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise B2IError(f"rc={rc}: {self.L.b2i_last_error(self.h).decode()}")

    def sync(self):
        self._check(self.L.b2i_ctx_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.L.b2i_ctx_launch_count(self.h))

    def decode_host(self, host_in, in_bytes: int, descs, host_out, out_bytes: int):
        """b2i_decode_host: host_in / host_out are ctypes buffers or integer addresses."""
        n = len(descs)
        res = (StreamResult * n)()
        self._check(self.L.b2i_decode_host(self.h, _addr(host_in), in_bytes, descs, n,
                                           _addr(host_out), out_bytes, res))
        return res

    def submit(self, host_in, in_bytes: int, descs, host_out, out_bytes: int):
        """b2i_submit: returns a job handle for wait(); at most two may be in flight."""
        job = C.c_void_p()
        self._check(self.L.b2i_submit(self.h, _addr(host_in), in_bytes, descs, len(descs),
                                      _addr(host_out), out_bytes, C.byref(job)))
        return (job, len(descs))

    def wait(self, job):
        """b2i_wait: blocks until the job's output is in host memory; returns its results."""
        handle, n = job
        res = (StreamResult * max(n, 1))()
        self._check(self.L.b2i_wait(handle, res))
        return res

    def crc32(self, data: bytes, crc: int = 0) -> int:
        out = C.c_uint32()
        self._check(self.L.b2i_crc32(self.h, crc, data if data is not None else None,
                                     len(data) if data is not None else 0, C.byref(out)))
        return out.value


def _addr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.cast(x, C.c_void_p)


def make_descs(items):
    arr = (StreamDesc * len(items))()
    for i, it in enumerate(items):
        arr[i] = it
    return arr


def zip_index(archive: bytes):
    """b2i_zip_index_build -> (list of dict, correction, has_encrypted)."""
    L = lib()
    ix = ZipIndex()
    err = C.create_string_buffer(128)
    rc = L.b2i_zip_index_build(archive, len(archive), C.byref(ix), err)
    if rc != OK:
        raise B2IError(f"zip index: {rc} {err.value.decode()}")
    names = C.string_at(ix.names, ix.names_len) if ix.names_len else b""
    out = []
    for i in range(ix.n):
        e = ix.entries[i]
        d = {f: getattr(e, f) for f, _ in ZipEntry._fields_}
        d["name"] = names[e.name_offset:e.name_offset + e.name_len]
        out.append(d)
    res = (out, ix.correction, bool(ix.has_encrypted_entries))
    L.b2i_zip_index_free(C.byref(ix))
    return res


def zip_index_via_fetch(archive: bytes, chunk: int = 1 << 16):
    """b2i_zip_index_build_cb over a source that only hands out bounded windows (what the
    libarchive plugin does for file-backed archives): -> (entries, correction, encrypted,
    bytes fetched)."""
    L = lib()
    ix = ZipIndex()
    err = C.create_string_buffer(128)
    keep = {"buf": None, "bytes": 0, "lo": 0, "hi": 0}

    def fetch(user, off, length):
        # a sliding read-ahead window, like the plugin keeps over its seekable source
        if not (keep["lo"] <= off and off + length <= keep["hi"]):
            n = max(length, min(chunk, len(archive) - off))
            keep["buf"] = C.create_string_buffer(archive[off:off + n], n)
            keep["lo"], keep["hi"] = off, off + n
            keep["bytes"] += n
        return C.addressof(keep["buf"]) + (off - keep["lo"])

    cb = FETCH_FN(fetch)
    rc = L.b2i_zip_index_build_cb(cb, None, len(archive), C.byref(ix), err)
    if rc != OK:
        raise B2IError(f"zip index: {rc} {err.value.decode()}")
    names = C.string_at(ix.names, ix.names_len) if ix.names_len else b""
    out = []
    for i in range(ix.n):
        e = ix.entries[i]
        d = {f: getattr(e, f) for f, _ in ZipEntry._fields_}
        d["name"] = names[e.name_offset:e.name_offset + e.name_len]
        out.append(d)
    res = (out, ix.correction, bool(ix.has_encrypted_entries), keep["bytes"])
    L.b2i_zip_index_free(C.byref(ix))
    return res


def gzip_scan_bgzf(buf: bytes, off: int = 0):
    L = lib()
    mem = C.POINTER(GzipMember)()
    n = C.c_size_t()
    end = C.c_size_t()
    rc = L.b2i_gzip_scan_bgzf(buf, len(buf), off, C.byref(mem), C.byref(n), C.byref(end))
    if rc != OK:
        raise B2IError(f"gzip scan: {rc}")
    out = [{f: getattr(mem[i], f) for f, _ in GzipMember._fields_} for i in range(n.value)]
    L.b2i_free(mem)
    return out, end.value


def partition_contiguous(descs, parts: int):
    """b2i_partition_contiguous -> [(lo, hi)] * parts (descriptor index ranges)."""
    cuts = (C.c_size_t * (parts + 1))()
    rc = lib().b2i_partition_contiguous(descs, len(descs), parts, cuts)
    if rc != OK:
        raise B2IError(f"b2i_partition_contiguous: {rc}")
    return [(int(cuts[i]), int(cuts[i + 1])) for i in range(parts)]


def partition_lpt(descs, parts: int):
    """b2i_partition_lpt -> (owner list, load list)."""
    n = len(descs)
    owner = (C.c_uint32 * max(n, 1))()
    load = (C.c_uint64 * parts)()
    rc = lib().b2i_partition_lpt(descs, n, parts, owner, load)
    if rc != OK:
        raise B2IError(f"b2i_partition_lpt: {rc}")
    return [int(owner[i]) for i in range(n)], [int(x) for x in load]


class Pipe:
    """b2i_pipe over an in-memory archive (bytes / ctypes buffer / address) or a fill callback."""

    def __init__(self, ctxs, descs, mem=None, mem_size=0, fill=None, window_out=0, first_window_out=0,
                 depth=0, copy_threads=0):
        self.L = lib()
        self.n = len(descs)
        self._keep = (mem, descs)
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        opts = PipeOpts(window_out, first_window_out, depth, copy_threads)
        self._cb = FILL_FN(fill) if fill is not None else C.cast(None, FILL_FN)
        h = C.c_void_p()
        rc = self.L.b2i_pipe_open(arr, len(ctxs), _addr(mem) if mem is not None else None, mem_size, self._cb,
                                  None, descs, self.n, C.byref(opts), C.byref(h))
        if rc != OK:
            raise B2IError(f"b2i_pipe_open: {rc}")
        self.h = h

    def get(self, idx):
        out, inp, res = C.c_void_p(), C.c_void_p(), StreamResult()
        rc = self.L.b2i_pipe_get(self.h, idx, C.byref(out), C.byref(inp), C.byref(res))
        if rc != OK:
            raise B2IError(f"b2i_pipe_get({idx}): {rc} {self.L.b2i_pipe_error(self.h).decode()}")
        return out.value, inp.value, res

    def release(self, idx):
        self.L.b2i_pipe_release(self.h, idx)

    @property
    def windows(self):
        return int(self.L.b2i_pipe_window_count(self.h))

    def close(self):
        if self.h:
            self.L.b2i_pipe_close(self.h)
            self.h = None
