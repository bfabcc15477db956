"""Host-side batch planning + a Python mirror of the reference read API for the
hot path, used by tests and bench.py (the C plugins in csrc/plugin/ are the
drop-in; this module exercises the same C ABI with the same semantics).

Names, return codes and messages follow the reference:
  archive_read_next_header / archive_read_data_block   archive_read.c:607-680, 966-982
  archive_read_format_zip_read_data                    archive_read_support_format_zip.c:3071-3198
  gzip_filter_read                                     archive_read_support_filter_gzip.c:431-511
"""
from __future__ import annotations

import ctypes as C

from . import capi
from .capi import StreamDesc

ARCHIVE_EOF, ARCHIVE_OK, ARCHIVE_WARN, ARCHIVE_FAILED, ARCHIVE_FATAL = 1, 0, -20, -25, -30
ZIP_BLOCK = 256 * 1024     # zip.c:2550: at most 256 KiB per read_data call
GZIP_BLOCK = 64 * 1024     # gzip.c:314: 64 KiB output block

_METHOD_NAMES = {0: "uncompressed", 1: "shrinking", 2: "reduced-1", 3: "reduced-2", 4: "reduced-3",
                 5: "reduced-4", 6: "imploded", 7: "reserved", 8: "deflation", 9: "deflation-64-bit",
                 10: "ibm-terse", 11: "reserved", 12: "bzip", 13: "reserved", 14: "lzma",
                 15: "reserved", 16: "reserved", 17: "reserved", 18: "ibm-terse-new", 19: "ibm-lz777",
                 93: "zstd", 95: "xz", 96: "jpeg", 97: "wav-pack", 98: "ppmd-1", 99: "aes"}


def compression_name(m: int) -> str:
    return _METHOD_NAMES.get(m, "??")


def _align16(v: int) -> int:
    return (v + 15) & ~15


def plan_zip(entries, ignore_crc32: bool = False, stored_no_copy: bool = True):
    """ZIP index -> (descs, out_bytes, which) : one descriptor per entry that has a
    body this build decodes (regular file, method 0/8, not encrypted, body present)."""
    items, which, out = [], [], 0
    for i, e in enumerate(entries):
        if e["warn"] & (capi_const.ZW_BAD_LOCAL_HEADER | capi_const.ZW_TRUNCATED):
            continue
        if (e["mode"] & 0o170000) != 0o100000:
            continue
        if e["zip_flags"] & 0x41 or e["method"] not in (0, 8) or e["compressed_size"] < 1:
            continue
        d = StreamDesc()
        d.in_off, d.in_len = e["data_offset"], e["compressed_size"]
        d.expect_out, d.expect_crc = e["uncompressed_size"], e["crc32"]
        d.method = e["method"]
        d.flags = (capi.F_NO_CRC if ignore_crc32 else 0)
        d.out_off = out
        if e["method"] == 0 and stored_no_copy:
            d.flags |= capi.F_NO_COPY
            d.out_cap = 0
        else:
            d.out_cap = e["uncompressed_size"] if e["method"] == 8 else e["compressed_size"]
            out = _align16(out + d.out_cap)
        items.append(d)
        which.append(i)
    return capi.make_descs(items), out, which


class capi_const:
    ZW_CRC, ZW_CSIZE, ZW_USIZE, ZW_BAD_LOCAL_HEADER, ZW_TRUNCATED = 1, 2, 4, 8, 16


class ZipReader:
    """archive_read_support_format_zip (seekable) over an in-memory archive."""

    def __init__(self, ctx: capi.Context, archive: bytes, options: str = ""):
        self.ctx = ctx
        self.archive = archive
        self.ignore_crc32 = "ignorecrc32" in options
        self.entries, self.correction, self.has_encrypted = capi.zip_index(archive)
        self.i = -1
        self.error = None
        self.format_name = "ZIP"
        self._decoded = False
        self._res = {}
        self._retry_bufs = {}

    # -- one device pass for the whole archive, on first use -----------------
    def _decode_all(self):
        self.descs, self.out_bytes, self.which = plan_zip(self.entries, self.ignore_crc32)
        n = len(self.descs)
        self.inbuf = C.create_string_buffer(self.archive, len(self.archive) + 32)
        self.outbuf = C.create_string_buffer(self.out_bytes + 32)
        if n:
            res = self.ctx.decode_host(self.inbuf, len(self.archive), self.descs, self.outbuf,
                                       self.out_bytes)
            for k, ei in enumerate(self.which):
                self._res[ei] = (self.descs[k], res[k])
            # a stream that outgrew its directory size is decoded again with room
            for k, ei in enumerate(self.which):
                cap = max(int(self.descs[k].out_cap), 1024)
                while self._res[ei][1].status == capi.S_OUT_OVERFLOW and cap < (1 << 31):
                    cap *= 4
                    d = StreamDesc.from_buffer_copy(self.descs[k])
                    d.out_off, d.out_cap = 0, cap
                    buf = C.create_string_buffer(cap + 32)
                    r = self.ctx.decode_host(self.inbuf, len(self.archive), capi.make_descs([d]), buf, cap)
                    self._res[ei] = (d, r[0])
                    self._retry_bufs[ei] = buf
        self._decoded = True

    def next_header(self):
        """-> (ARCHIVE_*, entry dict or None)"""
        self.i += 1
        if self.i >= len(self.entries):
            return ARCHIVE_EOF, None
        e = self.entries[self.i]
        if e["warn"] & capi_const.ZW_TRUNCATED and not e["name_len"]:
            self.error = "Truncated ZIP file header"
            return ARCHIVE_FATAL, None
        if e["warn"] & capi_const.ZW_BAD_LOCAL_HEADER:
            self.error = "Damaged Zip archive"
            return ARCHIVE_FATAL, None
        self.format_name = "ZIP %d.%d (%s)" % (e["version"] // 10, e["version"] % 10,
                                               compression_name(e["method"]))
        self._delivered = 0
        self._eof = e["compressed_size"] < 1
        ret = ARCHIVE_OK
        if e["warn"] & capi_const.ZW_CRC and not self.ignore_crc32:
            self.error, ret = "Inconsistent CRC32 values", ARCHIVE_WARN
        if e["warn"] & capi_const.ZW_CSIZE:
            ret = ARCHIVE_WARN
        if e["warn"] & capi_const.ZW_USIZE:
            ret = ARCHIVE_WARN
        return ret, e

    def read_data_block(self):
        """-> (ARCHIVE_*, bytes, offset) with the reference's block contract."""
        e = self.entries[self.i]
        off = self._delivered
        if self._eof or (e["mode"] & 0o170000) != 0o100000:
            return ARCHIVE_EOF, b"", off
        if e["zip_flags"] & 0x41:
            self.error = "Encrypted ZIP entries are not supported by this build"
            return ARCHIVE_FAILED, b"", off
        if e["method"] not in (0, 8):
            self.error = "Unsupported ZIP compression method (%d: %s)" % (
                e["method"], compression_name(e["method"]))
            return ARCHIVE_FAILED, b"", off
        if e["warn"] & capi_const.ZW_TRUNCATED:
            # the body runs past the end of the file: no descriptor was planned for it
            self.error = "Truncated ZIP file body"
            return ARCHIVE_FATAL, b"", off
        if not self._decoded:
            self._decode_all()
        if self.i not in self._res:
            self.error = "Truncated ZIP file body"
            return ARCHIVE_FATAL, b"", off
        d, r = self._res[self.i]
        total = int(r.out_bytes)
        if e["method"] == 0:
            # zero-copy from the archive, whole remaining body per call (open_memory)
            if off >= total:
                self._eof = True
                return self._final_checks(e, d, r, b"", off)
            blk = self.archive[e["data_offset"] + off:e["data_offset"] + total]
            self._delivered = total
            return ARCHIVE_OK, blk, off
        buf = self._retry_bufs.get(self.i, self.outbuf)
        n = min(ZIP_BLOCK, total - off)
        last = off + n >= total
        if r.status == capi.S_BUF_ERROR:
            # zlib hands out everything it could produce (Z_OK), then reports
            # Z_BUF_ERROR on the call that finds no input left (zip.c:2570-2657)
            if off >= total:
                self.error = "ZIP decompression failed (%d)" % r.status
                return ARCHIVE_FATAL, b"", off
            last = False
        elif r.status != capi.S_OK and n < ZIP_BLOCK:
            # the call that runs into the bad data fails; its partial output is dropped
            self.error = "ZIP decompression failed (%d)" % r.status
            return ARCHIVE_FATAL, b"", off
        elif r.status != capi.S_OK:
            last = False
        blk = buf.raw[d.out_off + off:d.out_off + off + n] if n else b""
        self._delivered = off + n
        if not last:
            return ARCHIVE_OK, blk, off
        self._eof = True
        return self._final_checks(e, d, r, blk, off)

    def _final_checks(self, e, d, r, blk, off):
        # order: CRC, compressed size, uncompressed size (zip.c:3164-3194)
        if r.flags & capi.R_CRC_MISMATCH and not self.ignore_crc32:
            self.error = "ZIP bad CRC: 0x%x should be 0x%x" % (r.crc, e["crc32"])
            return ARCHIVE_FAILED, b"", off
        if r.flags & capi.R_IN_MISMATCH:
            self.error = "ZIP compressed data is wrong size (read %d, expected %d)" % (
                r.in_bytes, e["compressed_size"])
            return ARCHIVE_FAILED, b"", off
        if r.flags & capi.R_OUT_MISMATCH:
            self.error = "ZIP uncompressed data is wrong size (read %d, expected %d)\n" % (
                r.out_bytes, e["uncompressed_size"])
            return ARCHIVE_FAILED, b"", off
        return ARCHIVE_OK, blk, off

    def read_data(self):
        """archive_read_data until EOF: -> (last status, all bytes)."""
        out = bytearray()
        while True:
            st, blk, _ = self.read_data_block()
            if st != ARCHIVE_OK:
                return st, bytes(out)
            out += blk


def plan_bgzf(members):
    items, out = [], 0
    for m in members:
        d = StreamDesc()
        d.in_off, d.in_len = m["deflate_offset"], m["deflate_len"]
        d.expect_out, d.expect_crc = m["isize"], m["crc32"]
        d.method = 8
        d.out_off, d.out_cap = out, m["isize"]
        out = _align16(out + m["isize"])
        items.append(d)
    return capi.make_descs(items), out


class GzipReader:
    """archive_read_support_filter_gzip + format raw over an in-memory file.

    BGZF members (BSIZE known) are decoded in one device pass; members without
    BSIZE are decoded one launch at a time, each launch locating the next
    header from the previous member's consumed byte count (still on the GPU)."""

    def __init__(self, ctx: capi.Context, data: bytes, verify_trailer: bool = True):
        self.ctx, self.data, self.verify = ctx, data, verify_trailer
        self.error = None
        self._out = None

    def bid(self) -> int:
        m = capi.GzipMember()
        return 27 if capi.lib().b2i_gzip_peek_header(self.data, len(self.data), 0, C.byref(m)) else 0

    def _decode(self):
        L = capi.lib()
        data, n = self.data, len(self.data)
        inbuf = C.create_string_buffer(data, n + 32)
        out = bytearray()
        off = 0
        status = ARCHIVE_OK
        while off < n:
            members, end = capi.gzip_scan_bgzf(data, off)
            if members:
                descs, out_bytes = plan_bgzf(members)
                obuf = C.create_string_buffer(out_bytes + 32)
                res = self.ctx.decode_host(inbuf, n, descs, obuf, out_bytes)
                for d, r in zip(descs, res):
                    if r.status != capi.S_OK or r.flags & capi.R_IN_MISMATCH:
                        self.error = "gzip decompression failed"
                        self._out, self._status = bytes(out), ARCHIVE_FATAL
                        return
                    if self.verify and r.flags & (capi.R_CRC_MISMATCH | capi.R_OUT_MISMATCH):
                        self.error = "gzip trailer CRC/ISIZE mismatch"
                        self._out, self._status = bytes(out), ARCHIVE_FATAL
                        return
                    out += obuf.raw[d.out_off:d.out_off + r.out_bytes]
                off = end
                continue
            m = capi.GzipMember()
            hl = L.b2i_gzip_peek_header(data, n, off, C.byref(m))
            if hl == 0:
                break               # trailing garbage: silent EOF (gzip.c:450-454)
            body = off + hl
            if body >= n:
                self.error, status = "truncated gzip input", ARCHIVE_FATAL
                break
            cap = 1 << 16
            while True:
                d = StreamDesc()
                d.in_off, d.in_len, d.method = body, n - body, 8
                d.out_cap, d.flags = cap, capi.F_NO_CRC if not self.verify else 0
                obuf = C.create_string_buffer(cap + 32)
                r = self.ctx.decode_host(inbuf, n, capi.make_descs([d]), obuf, cap)[0]
                if r.status != capi.S_OUT_OVERFLOW:
                    break
                cap *= 4
            if r.status == capi.S_BUF_ERROR:
                self.error, status = "truncated gzip input", ARCHIVE_FATAL
                break
            if r.status != capi.S_OK:
                self.error, status = "gzip decompression failed", ARCHIVE_FATAL
                break
            out += obuf.raw[:r.out_bytes]
            trailer = body + r.in_bytes
            if n - trailer < 8:
                status = ARCHIVE_FATAL          # consume_trailer: < 8 bytes (gzip.c:418-420)
                break
            if self.verify:
                crc = int.from_bytes(data[trailer:trailer + 4], "little")
                isz = int.from_bytes(data[trailer + 4:trailer + 8], "little")
                if crc != r.crc or isz != (r.out_bytes & 0xFFFFFFFF):
                    self.error, status = "gzip trailer CRC/ISIZE mismatch", ARCHIVE_FATAL
                    break
            off = trailer + 8
        self._out, self._status = bytes(out), status

    def read_all(self):
        """-> (status, bytes): ARCHIVE_EOF after all data, or ARCHIVE_FATAL."""
        if self._out is None:
            self._decode()
        return (ARCHIVE_EOF if self._status == ARCHIVE_OK else self._status), self._out
