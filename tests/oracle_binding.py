"""ctypes binding of oracle/liboracle.so and oracle/_ref/oracle_extract (the
checker).  Only tests/, smoke() and bench.py's cpu_baseline leg import this."""
import ctypes as C
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_EXTRACT = os.path.join(ROOT, "oracle", "_ref", "oracle_extract")


class OrcResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("detail", C.c_int32), ("out_bytes", C.c_uint64),
                ("in_bytes", C.c_uint64)]


class OrcStreamResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("crc", C.c_uint32), ("out_bytes", C.c_uint64),
                ("in_bytes", C.c_uint64), ("detail", C.c_uint32), ("flags", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                           stdout=subprocess.DEVNULL)
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_crc32.restype = C.c_uint32
        _lib.orc_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        _lib.orc_bitcrc32.restype = C.c_uint32
        _lib.orc_bitcrc32.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        _lib.orc_crc32_combine.restype = C.c_uint32
        _lib.orc_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        _lib.orc_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(OrcResult)]
        _lib.orc_decode_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_size_t, C.c_void_p]
    return _lib


def inflate(stream: bytes, cap: int):
    out = C.create_string_buffer(cap + 8)
    r = OrcResult()
    lib().orc_inflate(stream, len(stream), out, cap, C.byref(r))
    return r, out.raw[:r.out_bytes]


def decode_batch(inbytes: bytes, descs, out_bytes: int):
    """Oracle counterpart of b2i_decode_host (same descriptor layout)."""
    n = len(descs)
    res = (OrcStreamResult * n)()
    out = C.create_string_buffer(out_bytes + 32)
    rc = lib().orc_decode_batch(inbytes, len(inbytes), C.cast(descs, C.c_void_p), n, out, out_bytes,
                                C.cast(res, C.c_void_p))
    assert rc == 0
    return res, out.raw


def crc32(data: bytes, crc: int = 0) -> int:
    return lib().orc_crc32(crc, data, len(data))


def ref_list(path: str, raw: bool = False, opt: str | None = None, timeout: int = 60):
    """Run the UNMODIFIED reference (oracle/_ref) on a file; -> (report lines, data bytes)."""
    dump = path + ".refdump"
    cmd = [REF_EXTRACT, "list", path, "--dump", dump]
    if raw:
        cmd.append("--raw")
    if opt:
        cmd += ["--opt", opt]
    r = subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=timeout)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    with open(dump, "rb") as f:
        data = f.read()
    os.unlink(dump)
    return lines, data


def have_ref() -> bool:
    return os.path.exists(REF_EXTRACT)
