"""Throughput through libarchive's PUBLIC API (archive_read_open_memory, next_header,
archive_read_data_block / archive_read_data) of the drop-in library against the unmodified
reference, same process, same archive, one thread.  The first pass warms the device context."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(path):
    L = C.CDLL(path)
    L.archive_read_new.restype = C.c_void_p
    for f in ("archive_read_support_format_zip", "archive_read_support_filter_gzip", "archive_read_support_format_raw",
              "archive_read_free"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.archive_read_open_memory.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.archive_read_next_header.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.archive_read_data_block.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
    L.archive_read_data.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.archive_read_data.restype = C.c_ssize_t
    L.archive_error_string.argtypes = [C.c_void_p]
    L.archive_error_string.restype = C.c_char_p
    return L


def one_pass(L, blob, raw, copy):
    a = L.archive_read_new()
    if raw:
        L.archive_read_support_filter_gzip(a)
        L.archive_read_support_format_raw(a)
    else:
        L.archive_read_support_format_zip(a)
    buf = C.create_string_buffer(blob, len(blob))
    t0 = time.perf_counter()             # opening counts: the gzip filter decodes while formats bid
    assert L.archive_read_open_memory(a, buf, len(blob)) == 0, L.archive_error_string(a)
    e, p, n, o = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_int64()
    out = C.create_string_buffer(65536)
    total = 0
    while L.archive_read_next_header(a, C.byref(e)) == 0:
        if copy:
            while True:
                r = L.archive_read_data(a, out, 65536)
                if r <= 0:
                    assert r == 0, L.archive_error_string(a)
                    break
                total += r
        else:
            while True:
                r = L.archive_read_data_block(a, C.byref(p), C.byref(n), C.byref(o))
                if r != 0:
                    assert r == 1, L.archive_error_string(a)
                    break
                total += n.value
    dt = time.perf_counter() - t0
    L.archive_read_free(a)
    return total, dt


if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "zip64k"
    blob, kind = bench.build_workload(wl, 0, float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
    libs = [("reference", os.path.join(ROOT, "oracle", "_ref", "libarchive_ref.so")),
            ("drop-in", os.path.join(ROOT, "libarchive_b200", "libarchive_dropin.so"))]
    for name, path in libs:
        L = load(path)
        for copy in (False, True):
            best = None
            for rep in range(3):
                total, dt = one_pass(L, blob, kind == "bgzf", copy)
                best = dt if best is None or dt < best else best
            api = "archive_read_data(64 KiB)" if copy else "archive_read_data_block"
            print("%-9s %-26s %s: %d bytes, best of 3 %.1f ms = %.2f GB/s" % (name, api, wl, total, best * 1e3, total / best / 1e9), flush=True)
