"""EmulContext - TEST INFRASTRUCTURE ONLY: same interface as capi.Context.decode_host
but the per-stream device code runs under tests/emul (host SIMT emulation, one
warp = 32 pthreads).  Lets the not-gpu suite exercise the host-side reader logic
and the warp algorithms; it is slow and is never used by the product."""
import ctypes as C
import os
import subprocess

from libarchive_b200 import capi
from libarchive_b200.capi import StreamResult

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "emul", "libemul.so")
_emu = None
_libs = {}
VARIANT = ""          # "" = default build, "r9" = the 9-bit-root build of the inflate kernel


def emu():
    global _emu
    if VARIANT not in _libs:
        subprocess.run(["make", "-C", os.path.join(HERE, "emul")], check=True, stdout=subprocess.DEVNULL)
        lib = C.CDLL(SO if not VARIANT else SO.replace("libemul.so", "libemul_%s.so" % VARIANT))
        lib.emul_crc32.restype = C.c_uint32
        lib.emul_crc32.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64]
        lib.emul_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        _libs[VARIANT] = lib
    _emu = _libs[VARIANT]
    return _emu


class EmulContext:
    launch_count = 0

    def decode_host(self, host_in, in_bytes, descs, host_out, out_bytes):
        n = len(descs)
        res = (StreamResult * n)()
        E = emu()
        for k in range(n):
            d, r = descs[k], res[k]
            if d.method == 8:
                E.emul_inflate(host_in, in_bytes, host_out, C.byref(d), C.byref(r))
            elif d.method == 0:
                r.out_bytes = r.in_bytes = d.in_len
                if not d.flags & capi.F_NO_CRC:
                    r.crc = E.emul_crc32(0, host_in, d.in_off, d.in_len)
                    if r.crc != d.expect_crc:
                        r.flags |= capi.R_CRC_MISMATCH
                if (d.in_len & 0xFFFFFFFF) != (d.expect_out & 0xFFFFFFFF):
                    r.flags |= capi.R_OUT_MISMATCH
                if not d.flags & capi.F_NO_COPY:
                    C.memmove(C.addressof(host_out) + d.out_off, C.addressof(host_in) + d.in_off, d.in_len)
            else:
                r.status = capi.S_UNSUPPORTED
        return res

    def crc32(self, data, crc=0):
        if data is None:
            return 0
        buf = C.create_string_buffer(data, len(data) + 16)
        return emu().emul_crc32(crc, buf, 0, len(data))

    def close(self):
        pass
