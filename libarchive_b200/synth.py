"""Deterministic synthetic inputs for the five BASELINE.json configs.

Everything here is input *construction* (writers, generators); nothing decodes.
Compression uses Python's zlib (the same system zlib the reference links), with
the parameters the reference's ZIP writer uses:
``deflateInit2(level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY)``
(archive_write_set_format_zip.c:1339-1340).

Containers are written by hand from the format specifications so that both
framings the reader must accept can be produced:
  * "sizes"   - CRC and sizes in the local header (what `zipfile`/Info-ZIP do)
  * "at_end"  - length-at-end flag, zero CRC/sizes in the local header and a
                data descriptor after the data, which is what the reference's
                own writer always emits (archive_write_set_format_zip.c:1052,
                1131-1135)
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

# ---------------------------------------------------------------------------
# payload generators
# ---------------------------------------------------------------------------


def _vocabulary(seed: int = 12345, nwords: int = 4096):
    rng = np.random.default_rng(seed)
    lens = rng.integers(2, 11, size=nwords)  # 2..10 letters
    table = np.full((nwords, 11), ord(" "), dtype=np.uint8)
    for i, n in enumerate(lens):
        table[i, :n] = rng.integers(ord("a"), ord("z") + 1, size=n)
    weights = 1.0 / np.arange(1, nwords + 1)  # Zipf(1)
    weights /= weights.sum()
    return table, lens.astype(np.int64) + 1, weights  # +1: trailing separator


def synth_text(nbytes: int, seed: int = 12345) -> bytes:
    """'Synthetic text' of SURVEY section 8(d): 4096 pseudo-words (2-10 lowercase
    letters), Zipf(1) frequencies, space separated, newline every 2048 words."""
    table, wlens, weights = _vocabulary()
    rng = np.random.default_rng(seed)
    out = bytearray()
    mean = float((wlens * weights).sum())
    while len(out) < nbytes:
        n = int((nbytes - len(out)) / mean) + 64
        idx = rng.choice(len(weights), size=n, p=weights)
        lens = wlens[idx]
        ends = np.cumsum(lens)
        starts = ends - lens
        total = int(ends[-1])
        word_of = np.repeat(np.arange(n), lens)
        col = np.arange(total) - np.repeat(starts, lens)
        buf = table[idx[word_of], col]
        buf[ends[2047::2048] - 1] = ord("\n")
        out += buf.tobytes()
    return bytes(out[:nbytes])


def synth_random(nbytes: int, seed: int = 1) -> bytes:
    return np.random.default_rng(seed).integers(0, 256, size=nbytes, dtype=np.uint8).tobytes()


def deflate_raw(data: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
                mem_level: int = 8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    return c.compress(data) + c.flush()


def deflate_mixed(parts) -> bytes:
    """One raw deflate stream whose blocks change type: `parts` is a list of
    (bytes, level, strategy); parts are separated by Z_FULL_FLUSH (an empty
    stored block) so dynamic, fixed and stored blocks interleave."""
    out = bytearray()
    # a single compressobj cannot change strategy mid-stream portably, so each
    # part is its own stream with BFINAL patched off except for the last one.
    for k, (data, level, strategy) in enumerate(parts):
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        body = c.compress(data) + c.flush(zlib.Z_FULL_FLUSH)
        out += body
    out += b"\x03\x00"  # final empty fixed-Huffman block (BFINAL=1, BTYPE=01, EOB)
    return bytes(out)


# ---------------------------------------------------------------------------
# ZIP writer
# ---------------------------------------------------------------------------


@dataclass
class ZipMember:
    name: str
    data: bytes                 # uncompressed payload
    method: int = 8             # 0 stored, 8 deflate
    level: int = 6
    strategy: int = zlib.Z_DEFAULT_STRATEGY
    comp: bytes | None = None   # pre-compressed payload (overrides data/level)
    crc: int | None = None      # override (to plant a wrong CRC)
    usize: int | None = None    # override uncompressed size field
    csize: int | None = None    # override compressed size field
    flags: int = 0


def _compress_member(m: ZipMember) -> bytes:
    if m.comp is not None:
        return m.comp
    if m.method == 0:
        return m.data
    return deflate_raw(m.data, m.level, m.strategy)


def make_zip(members, framing: str = "sizes", zip64: bool = False, prefix: bytes = b"",
             threads: int = 8, comment: bytes = b"") -> bytes:
    """Write a ZIP archive.  framing: 'sizes' | 'at_end' (see module doc)."""
    members = list(members)
    if threads > 1 and len(members) > 16:
        with ThreadPoolExecutor(threads) as ex:
            comps = list(ex.map(_compress_member, members, chunksize=64))
    else:
        comps = [_compress_member(m) for m in members]
    out = bytearray(prefix)
    cd = bytearray()
    force64 = zip64
    for m, comp in zip(members, comps):
        name = m.name.encode("utf-8")
        crc = zlib.crc32(m.data) & 0xFFFFFFFF if m.crc is None else m.crc
        usize = len(m.data) if m.usize is None else m.usize
        csize = len(comp) if m.csize is None else m.csize
        off = len(out) - len(prefix)
        flags = m.flags | (0x08 if framing == "at_end" else 0)
        need64 = force64 or usize >= 0xFFFFFFFF or csize >= 0xFFFFFFFF
        ver = 45 if need64 else (20 if m.method == 8 else 10)
        lextra = b""
        if framing == "at_end":
            lcrc = lusz = lcsz = 0
        else:
            lcrc, lusz, lcsz = crc, usize, csize
            if need64:
                lextra = struct.pack("<HHQQ", 1, 16, usize, csize)
                lusz = lcsz = 0xFFFFFFFF
        out += struct.pack("<4sHHHIIIIHH", b"PK\x03\x04", ver, flags, m.method, 0x58C55A2B & 0xFFFFFFFF,
                           lcrc, lcsz, lusz, len(name), len(lextra))
        out += name + lextra + comp
        if framing == "at_end":
            if need64:
                out += struct.pack("<4sIQQ", b"PK\x07\x08", crc, csize, usize)
            else:
                out += struct.pack("<4sIII", b"PK\x07\x08", crc, csize, usize)
        cextra = b""
        c_usz, c_csz, c_off = usize, csize, off
        if need64 or off >= 0xFFFFFFFF:
            fields = b""
            if need64 or usize >= 0xFFFFFFFF:
                fields += struct.pack("<Q", usize); c_usz = 0xFFFFFFFF
            if need64 or csize >= 0xFFFFFFFF:
                fields += struct.pack("<Q", csize); c_csz = 0xFFFFFFFF
            if off >= 0xFFFFFFFF:
                fields += struct.pack("<Q", off); c_off = 0xFFFFFFFF
            cextra = struct.pack("<HH", 1, len(fields)) + fields
        cd += struct.pack("<4sBBHHHIIIIHHHHHII", b"PK\x01\x02", ver, 3, ver, flags, m.method,
                          0x58C55A2B & 0xFFFFFFFF, crc, c_csz, c_usz, len(name), len(cextra), 0, 0, 0,
                          (0o100644 << 16) & 0xFFFFFFFF, c_off)
        cd += name + cextra
    cd_off = len(out) - len(prefix)
    out += cd
    n = len(members)
    if force64 or n > 0xFFFF or cd_off >= 0xFFFFFFFF or len(cd) >= 0xFFFFFFFF:
        eocd64_off = len(out) - len(prefix)
        out += struct.pack("<4sQHHIIQQQQ", b"PK\x06\x06", 44, 45, 45, 0, 0, n, n, len(cd), cd_off)
        out += struct.pack("<4sIQI", b"PK\x06\x07", 0, eocd64_off, 1)
        out += struct.pack("<4sHHHHIIH", b"PK\x05\x06", 0, 0, min(n, 0xFFFF), min(n, 0xFFFF),
                           min(len(cd), 0xFFFFFFFF), min(cd_off, 0xFFFFFFFF), len(comment))
    else:
        out += struct.pack("<4sHHHHIIH", b"PK\x05\x06", 0, 0, n, n, len(cd), cd_off, len(comment))
    out += comment
    return bytes(out)


# ---------------------------------------------------------------------------
# gzip / BGZF writer
# ---------------------------------------------------------------------------

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_member(data: bytes, level: int = 6) -> bytes:
    comp = deflate_raw(data, level)
    bsize = len(comp) + 25  # 18 header + 8 trailer - 1
    if bsize > 0xFFFF:
        raise ValueError("BGZF member too large")
    hdr = struct.pack("<4BIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, ord("B"), ord("C"), 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


def make_bgzf(blocks, level: int = 6, eof: bool = True, threads: int = 8) -> bytes:
    blocks = list(blocks)
    if threads > 1 and len(blocks) > 16:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda b: bgzf_member(b, level), blocks, chunksize=64))
    else:
        parts = [bgzf_member(b, level) for b in blocks]
    return b"".join(parts) + (BGZF_EOF if eof else b"")


def gzip_member(data: bytes, level: int = 6, name: bytes | None = None, comment: bytes | None = None,
                extra: bytes | None = None, hcrc: bool = False, mtime: int = 0) -> bytes:
    flg = (4 if extra is not None else 0) | (8 if name is not None else 0) | \
          (16 if comment is not None else 0) | (2 if hcrc else 0)
    h = struct.pack("<4BIBB", 0x1F, 0x8B, 8, flg, mtime, 0, 3)
    if extra is not None:
        h += struct.pack("<H", len(extra)) + extra
    if name is not None:
        h += name + b"\0"
    if comment is not None:
        h += comment + b"\0"
    if hcrc:
        h += struct.pack("<H", zlib.crc32(h) & 0xFFFF)
    return h + deflate_raw(data, level) + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF,
                                                      len(data) & 0xFFFFFFFF)


# ---------------------------------------------------------------------------
# the five BASELINE.json configs (scale < 1 shrinks counts for tests)
# ---------------------------------------------------------------------------


def split_text(total: int, each: int, seed: int):
    blob = synth_text(total, seed)
    return [blob[i:i + each] for i in range(0, total, each)]


def text_corpus_parts(total: int, each: int, seed: int, base: int = 64 << 20):
    """`total` bytes of synthetic text in pieces of `each` bytes for the LARGE configs: the
    generator above runs at ~15 MB/s, so beyond `base` bytes the pieces are cut from
    rotations of one `base`-byte text (rotation j starts 4099*j bytes in): same statistics,
    every piece a different byte string."""
    if total <= base:
        return split_text(total, each, seed)
    blob = synth_text(base, seed)
    parts, rot = [], 0
    while len(parts) * each < total:
        r = (4099 * rot) % (base - each)
        view = blob[r:] + blob[:r]
        for i in range(0, base - each + 1, each):
            parts.append(view[i:i + each])
            if len(parts) * each >= total:
                break
        rot += 1
    return parts


def config1_zip_text64k(n_entries: int = 4096, entry: int = 65536, framing: str = "sizes",
                        seed: int = 12345) -> bytes:
    parts = split_text(n_entries * entry, entry, seed)
    return make_zip([ZipMember("e%06d.txt" % i, p) for i, p in enumerate(parts)], framing=framing)


def config2_zip_stored1m(n_entries: int = 1024, entry: int = 1 << 20, seed: int = 2) -> bytes:
    blob = synth_random(n_entries * entry, seed)
    return make_zip([ZipMember("s%05d.bin" % i, blob[i * entry:(i + 1) * entry], method=0)
                     for i in range(n_entries)])


def config3_bgzf_text64k(n_members: int = 65536, member: int = 65536 - 512, seed: int = 54321) -> bytes:
    # every member must fit BSIZE (u16) even if incompressible, hence < 64 KiB of
    # text would never be a problem at 2.9:1; the SURVEY shape is 65536-byte members
    parts = split_text(n_members * member, member, seed)
    return make_bgzf(parts)


def config4_sizes(total: int = 8 << 30, lo: int = 1 << 10, hi: int = 16 << 20, seed: int = 4):
    rng = np.random.default_rng(seed)
    sizes, acc = [], 0
    while acc < total:
        s = int(round(lo * (hi / lo) ** rng.random()))
        s = min(s, total - acc)
        sizes.append(s)
        acc += s
    return sizes


def config4_zip64_mixed(total: int = 8 << 30, lo: int = 1 << 10, hi: int = 16 << 20, seed: int = 4,
                        force=(), threads: int = 8) -> bytes:
    """BASELINE config 4.  `force` = extra (size, kind) entries put in front of the
    log-uniform ones (kind 0 dynamic, 3 fixed, 6 stored blocks, 9 interleaved), so that a
    scaled-down archive still holds entries of the maximum size."""
    sizes = [int(s) for s, _ in force] + config4_sizes(total, lo, hi, seed)
    rng = np.random.default_rng(seed + 1)
    text = synth_text(min(max(sizes) + (1 << 20), 64 << 20), seed)
    members = []
    for i, s in enumerate(sizes):
        kind = int(rng.integers(0, 10))
        if i < len(force):
            kind = int(force[i][1])
        o = int(rng.integers(0, max(1, len(text) - s))) if s <= len(text) else 0
        payload = (text * (s // len(text) + 1))[o:o + s] if s > len(text) - o else text[o:o + s]
        if kind < 3:      # dynamic Huffman
            members.append(ZipMember("m%05d.txt" % i, payload, level=6))
        elif kind < 6:    # fixed Huffman
            members.append(ZipMember("m%05d.fix" % i, payload, level=1, strategy=zlib.Z_FIXED))
        elif kind < 9:    # random bytes: zlib emits stored blocks
            members.append(ZipMember("m%05d.rnd" % i, synth_random(s, seed + i)))
        else:             # all three interleaved in one stream
            a, b = s // 3, 2 * s // 3
            comp = deflate_mixed([(payload[:a], 6, zlib.Z_DEFAULT_STRATEGY),
                                  (synth_random(b - a, seed + i), 6, zlib.Z_DEFAULT_STRATEGY),
                                  (payload[b:], 1, zlib.Z_FIXED)])
            data = payload[:a] + synth_random(b - a, seed + i) + payload[b:]
            members.append(ZipMember("m%05d.mix" % i, data, comp=comp))
    return make_zip(members, zip64=True, threads=threads)


def config5_zip64_tiny(n_entries: int = 500_000, entry: int = 4096, seed: int = 5) -> bytes:
    parts = split_text(n_entries * entry, entry, seed)
    return make_zip([ZipMember("t%06d" % i, p) for i, p in enumerate(parts)], zip64=True)


# ---------------------------------------------------------------------------
# hand-assembled deflate streams (block-type zoo, malformed cases)
# ---------------------------------------------------------------------------


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value: int, n: int):          # LSB first (header fields, extra bits)
        for i in range(n):
            self.bits.append((value >> i) & 1)
        return self

    def code(self, code: int, n: int):          # Huffman codes go MSB first
        for i in range(n - 1, -1, -1):
            self.bits.append((code >> i) & 1)
        return self

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)
        return self

    def raw(self, data: bytes):
        for b in data:
            self.put(b, 8)
        return self

    def bytes(self) -> bytes:
        bits = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(sum(bits[i + k] << k for k in range(8)) for i in range(0, len(bits), 8))


def fixed_lit(w: BitWriter, sym: int):
    if sym < 144:
        w.code(0x30 + sym, 8)
    elif sym < 256:
        w.code(0x190 + sym - 144, 9)
    elif sym < 280:
        w.code(sym - 256, 7)
    else:
        w.code(0xC0 + sym - 280, 8)
    return w


_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
             131, 163, 195, 227, 258]
_LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
              2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577]
_DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12,
               13, 13]


def fixed_match(w: BitWriter, length: int, dist: int):
    ls = max(i for i in range(29) if _LEN_BASE[i] <= length and (i < 28 or length == 258))
    if length == 258:
        ls = 28
    fixed_lit(w, 257 + ls)
    w.put(length - _LEN_BASE[ls], _LEN_EXTRA[ls])
    ds = max(i for i in range(30) if _DIST_BASE[i] <= dist)
    w.code(ds, 5)
    w.put(dist - _DIST_BASE[ds], _DIST_EXTRA[ds])
    return w


def dynamic_header(w: BitWriter, litlen_lens, dist_lens, final: int = 1, hclen_all: bool = True):
    """Emit BFINAL/BTYPE=2 and a header coding the given length arrays with a
    trivial code-length code (every code length symbol 0..15 gets 4 bits...
    i.e. a flat 4-bit code over symbols 0..15, no repeats)."""
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    w.put(final, 1).put(2, 2)
    w.put(len(litlen_lens) - 257, 5).put(len(dist_lens) - 1, 5).put(19 - 4, 4)
    cl = [0] * 19
    for s in range(16):
        cl[s] = 4
    for s in order:
        w.put(cl[s], 3)
    # flat 4-bit canonical code: symbol s has code s
    for l in list(litlen_lens) + list(dist_lens):
        w.code(l, 4)
    return w


def canonical_codes(lens):
    maxl = max(lens) if lens else 0
    count = [0] * (maxl + 2)
    for l in lens:
        if l:
            count[l] += 1
    code, nxt = 0, [0] * (maxl + 2)
    for bits in range(1, maxl + 1):
        code = (code + count[bits - 1]) << 1
        nxt[bits] = code
    out = []
    for l in lens:
        if l:
            out.append(nxt[l]); nxt[l] += 1
        else:
            out.append(None)
    return out


def deflate_zoo():
    """Named raw-deflate streams covering the zlib acceptance rules of SURVEY
    section 8(c).  Returns list of (name, bytes).  Expected results come from
    Python's zlib in the tests, never from a table here."""
    zoo = []

    def add(name, w):
        zoo.append((name, w.bytes() if isinstance(w, BitWriter) else bytes(w)))

    # stored blocks
    add("stored_empty_final", BitWriter().put(1, 1).put(0, 2).align().put(0, 16).put(0xFFFF, 16))
    add("stored_hello", BitWriter().put(1, 1).put(0, 2).align().put(5, 16).put(0xFFFA, 16).raw(b"hello"))
    add("stored_two_blocks", BitWriter().put(0, 1).put(0, 2).align().put(3, 16).put(0xFFFC, 16).raw(b"abc")
        .put(1, 1).put(0, 2).align().put(2, 16).put(0xFFFD, 16).raw(b"de"))
    add("stored_bad_nlen", BitWriter().put(1, 1).put(0, 2).align().put(5, 16).put(0x1234, 16).raw(b"hello"))
    add("stored_truncated", BitWriter().put(1, 1).put(0, 2).align().put(5, 16).put(0xFFFA, 16).raw(b"he"))
    add("stored_max", BitWriter().put(1, 1).put(0, 2).align().put(65535, 16).put(0, 16)
        .raw(bytes(range(256)) * 255 + bytes(range(255))))
    add("block_type_3", BitWriter().put(1, 1).put(3, 2).put(0, 13))
    add("empty_input", b"")
    add("one_byte_only_header", BitWriter().put(0, 1).put(1, 2))
    # fixed blocks
    w = BitWriter().put(1, 1).put(1, 2)
    for c in b"hello\n":
        fixed_lit(w, c)
    fixed_match(w, 12, 6)
    fixed_lit(w, 256)
    add("fixed_hello_x3", w)
    add("fixed_empty", fixed_lit(BitWriter().put(1, 1).put(1, 2), 256))
    w = BitWriter().put(1, 1).put(1, 2)
    fixed_lit(w, ord("a"))
    fixed_match(w, 258, 1)
    fixed_match(w, 258, 1)
    fixed_lit(w, 256)
    add("fixed_rle_dist1_len258", w)
    w = BitWriter().put(1, 1).put(1, 2)
    for c in b"abc":
        fixed_lit(w, c)
    fixed_match(w, 3, 4)
    fixed_lit(w, 256)
    add("fixed_dist_too_far", w)
    w = BitWriter().put(1, 1).put(1, 2)
    fixed_lit(w, ord("a"))
    fixed_lit(w, 286)
    add("fixed_litlen_286", w)
    w = BitWriter().put(1, 1).put(1, 2)
    fixed_lit(w, ord("a"))
    fixed_lit(w, 257)
    w.code(30, 5)
    add("fixed_dist_code_30", w)
    w = BitWriter().put(1, 1).put(1, 2)
    for c in b"no end of block":
        fixed_lit(w, c)
    add("fixed_truncated_no_eob", w)
    # overlapping copies with every small distance
    w = BitWriter().put(1, 1).put(1, 2)
    for c in bytes(range(65, 65 + 40)):
        fixed_lit(w, c)
    for d in list(range(1, 41)):
        fixed_match(w, 3 + (d * 7) % 200, d)
    fixed_lit(w, 256)
    add("fixed_overlap_all_small_dists", w)
    # max distance
    w = BitWriter().put(0, 1).put(0, 2).align().put(32768, 16).put(32768 ^ 0xFFFF, 16)
    w.raw(synth_random(32768, 7))
    w.put(1, 1).put(1, 2)
    fixed_match(w, 258, 32768)
    fixed_match(w, 100, 32768)
    fixed_lit(w, 256)
    add("stored_then_fixed_maxdist", w)

    # dynamic: literal-only tree with a single distance code of length 1 (incomplete, accepted)
    ll = [0] * 257
    ll[ord("a")] = 1
    ll[256] = 1
    w = dynamic_header(BitWriter(), ll, [1])
    cc = canonical_codes(ll)
    w.code(cc[ord("a")], 1).code(cc[ord("a")], 1).code(cc[256], 1)
    add("dyn_single_dist_code", w)
    # dynamic: zero distance codes, literals only
    w = dynamic_header(BitWriter(), ll, [0])
    w.code(cc[ord("a")], 1).code(cc[256], 1)
    add("dyn_no_dist_codes_literals_only", w)
    # dynamic: zero distance codes but a length symbol is used -> invalid distance code
    ll2 = [0] * 258
    ll2[ord("a")] = 2
    ll2[256] = 2
    ll2[257] = 1
    cc2 = canonical_codes(ll2)
    w = dynamic_header(BitWriter(), ll2, [0])
    w.code(cc2[ord("a")], 2).code(cc2[257], 1).put(0, 1)
    add("dyn_no_dist_codes_but_match", w)
    # incomplete lit/len tree with max length != 1
    ll3 = [0] * 257
    ll3[ord("a")] = 2
    ll3[256] = 2
    add("dyn_incomplete_litlen", dynamic_header(BitWriter(), ll3, [1]).put(0, 16))
    # only EOB with length 1 (incomplete, max = 1): accepted, empty output
    ll4 = [0] * 257
    ll4[256] = 1
    w = dynamic_header(BitWriter(), ll4, [0])
    w.code(0, 1)
    add("dyn_only_eob", w)
    # ... and the unused 1-bit code is an invalid literal/length code
    w = dynamic_header(BitWriter(), ll4, [0])
    w.code(1, 1)
    add("dyn_only_eob_invalid_code", w)
    # over-subscribed
    ll5 = [0] * 257
    ll5[0] = ll5[1] = ll5[2] = 1
    ll5[256] = 1
    add("dyn_oversubscribed_litlen", dynamic_header(BitWriter(), ll5, [1]).put(0, 16))
    # over-subscribed / incomplete distance trees
    add("dyn_oversubscribed_dist", dynamic_header(BitWriter(), ll, [1, 1, 1]).put(0, 16))
    add("dyn_incomplete_dist", dynamic_header(BitWriter(), ll, [2, 2, 2]).put(0, 16))
    # missing end-of-block code
    ll6 = [0] * 257
    ll6[0] = ll6[1] = 1
    add("dyn_missing_eob", dynamic_header(BitWriter(), ll6, [1]).put(0, 16))
    # HLIT > 286
    w = BitWriter().put(1, 1).put(2, 2).put(30, 5).put(0, 5).put(0, 4).put(0, 64)
    add("dyn_hlit_287", w)
    # HDIST > 30
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(30, 5).put(0, 4).put(0, 64)
    add("dyn_hdist_31", w)
    # all-zero code-length code (zlib quirk: lengths read as 0 -> missing EOB)
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(0, 4).put(0, 12).put(0, 400)
    add("dyn_all_zero_codelen_code", w)
    # the same but truncated before 258 one-bit lengths are available
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(0, 4).put(0, 12).put(0, 100)
    add("dyn_all_zero_codelen_code_short", w)
    # repeat code 16 with no previous length
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(0, 4)
    order_vals = {16: 1, 17: 0, 18: 0, 0: 1}
    for s in [16, 17, 18, 0]:
        w.put(order_vals[s], 3)
    w.code(1, 1).put(0, 2).put(0, 32)   # symbol 16 (code '1'), 2 extra bits
    add("dyn_repeat_without_previous", w)
    # repeat overflowing the table
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(0, 4)
    for s, v in [(16, 0), (17, 0), (18, 1), (0, 1)]:
        w.put(v, 3)
    w.code(1, 1).put(127, 7).code(1, 1).put(127, 7).put(0, 32)
    add("dyn_repeat_overflow", w)
    # incomplete code-length code
    w = BitWriter().put(1, 1).put(2, 2).put(0, 5).put(0, 5).put(0, 4)
    for s, v in [(16, 0), (17, 0), (18, 0), (0, 1)]:
        w.put(v, 3)
    w.put(0, 64)
    add("dyn_incomplete_codelen_code", w)
    # long codes: 286 lit/len symbols with lengths 1..6,14,15 and a distance code
    # with lengths 1..15 (exercises the secondary lookup tables)
    ll7 = [15] * 286
    for k, sym in enumerate([ord("e"), ord(" "), ord("t"), 256, 257, ord("a")]):
        ll7[sym] = k + 1
    left = [i for i in range(286) if ll7[i] == 15]
    for sym in left[:232]:
        ll7[sym] = 14
    dl7 = [0] * 30
    for k in range(14):
        dl7[k] = k + 1
    dl7[14] = dl7[29] = 15
    zoo.append(("dyn_long_codes", encode_dynamic(ll7, dl7, _zoo_symbols())))
    return zoo


def _zoo_symbols():
    """literal / (length, distance) sequence touching every literal, every
    length symbol and the distance symbols 0..14 and 29"""
    seq = [("lit", b) for b in range(256)] * 2
    seq += [("lit", b) for b in b"the quick brown fox jumps over the lazy dog " * 40]
    for ls in range(29):
        length = _LEN_BASE[ls] + ((1 << _LEN_EXTRA[ls]) - 1 if _LEN_EXTRA[ls] else 0)
        for ds in list(range(15)):
            dist = _DIST_BASE[ds] + ((1 << _DIST_EXTRA[ds]) - 1) // 2
            seq.append(("match", length, dist))
    seq += [("lit", b) for b in synth_random(30000, 99)]
    seq.append(("match", 258, 24577 + 100))
    seq.append(("match", 3, 1))
    return seq


def encode_dynamic(litlen_lens, dist_lens, seq, final: int = 1) -> bytes:
    """Encode a symbol sequence with explicitly chosen code lengths."""
    w = dynamic_header(BitWriter(), litlen_lens, dist_lens, final)
    lc = canonical_codes(litlen_lens)
    dc = canonical_codes(dist_lens)
    for item in seq:
        if item[0] == "lit":
            w.code(lc[item[1]], litlen_lens[item[1]])
        else:
            _, length, dist = item
            ls = 28 if length == 258 else max(i for i in range(28) if _LEN_BASE[i] <= length)
            w.code(lc[257 + ls], litlen_lens[257 + ls])
            w.put(length - _LEN_BASE[ls], _LEN_EXTRA[ls])
            ds = max(i for i in range(30) if _DIST_BASE[i] <= dist)
            w.code(dc[ds], dist_lens[ds])
            w.put(dist - _DIST_BASE[ds], _DIST_EXTRA[ds])
    w.code(lc[256], litlen_lens[256])
    return w.bytes()


def random_code_lengths(rng, nsyms: int, nused: int, maxlen: int = 15):
    """Random COMPLETE prefix code over `nused` of `nsyms` symbols (random
    splitting of the Kraft tree), lengths <= maxlen."""
    leaves = [0]
    while len(leaves) < nused:
        cand = [i for i, d in enumerate(leaves) if d < maxlen]
        i = cand[int(rng.integers(0, len(cand)))] if rng.random() < 0.5 else \
            max(cand, key=lambda k: leaves[k] + rng.random())
        d = leaves.pop(i)
        leaves += [d + 1, d + 1]
    lens = [0] * nsyms
    syms = rng.permutation(nsyms)[:nused]
    for s, d in zip(syms, leaves):
        lens[int(s)] = max(d, 1)
    return lens


def random_dynamic_stream(seed: int, nsym: int = 3000) -> bytes:
    """A single dynamic block with random complete codes (often with 11..15-bit
    codes) and a random valid symbol sequence."""
    rng = np.random.default_rng(seed)
    while True:
        ll = random_code_lengths(rng, 286, int(rng.integers(2, 287)))
        if ll[256] == 0:
            # make sure EOB is coded: swap with some used symbol
            used = [i for i, l in enumerate(ll) if l]
            j = used[int(rng.integers(0, len(used)))]
            ll[256], ll[j] = ll[j], 0
        nd = int(rng.integers(1, 31))
        dl = random_code_lengths(rng, 30, nd) if nd > 1 else [1] + [0] * 29
        if len([l for l in ll if l]) >= 2:
            break
    lits = [i for i in range(256) if ll[i]]
    lsyms = [i for i in range(257, 286) if ll[i]]
    dsyms = [i for i in range(30) if dl[i]]
    seq, produced = [], 0
    for _ in range(nsym):
        if lsyms and dsyms and produced > 0 and (not lits or rng.random() < 0.4):
            ls = lsyms[int(rng.integers(0, len(lsyms)))] - 257
            length = _LEN_BASE[ls] + int(rng.integers(0, 1 << _LEN_EXTRA[ls]))
            if length == 258 and ls != 28:
                length = 257
            ok = [d for d in dsyms if _DIST_BASE[d] <= produced]
            if not ok:
                continue
            ds = ok[int(rng.integers(0, len(ok)))]
            dist = min(_DIST_BASE[ds] + int(rng.integers(0, 1 << _DIST_EXTRA[ds])), produced)
            if max(i for i in range(30) if _DIST_BASE[i] <= dist) != ds:
                continue
            seq.append(("match", length, dist))
            produced += length
        elif lits:
            seq.append(("lit", lits[int(rng.integers(0, len(lits)))]))
            produced += 1
    return encode_dynamic(ll, dl, seq)
