"""Generated archives and the comparison rule for the STREAMING ZIP reader (non-seekable
input), shared by the no-GPU test (plugin host logic over the oracle shim) and the GPU test
(the real drop-in).  Both are compared with the unmodified reference driven the same way."""
import json
import os
import subprocess
import tempfile
import zlib

from libarchive_b200 import synth


def archive(framing):
    txt = synth.synth_text(300000, 11)
    comp = synth.deflate_raw(txt, 6)
    crc = zlib.crc32(txt) & 0xFFFFFFFF
    members = [synth.ZipMember("dir/", b"", method=0), synth.ZipMember("dir/ok.txt", txt),
               synth.ZipMember("stored.bin", synth.synth_random(70000, 3), method=0),
               synth.ZipMember("empty", b"", method=0),
               synth.ZipMember("fixed.txt", txt[:5000], level=1, strategy=zlib.Z_FIXED),
               synth.ZipMember("crc.txt", txt[:40000], crc=crc ^ 1), synth.ZipMember("last.txt", txt[:777])]
    if framing == "sizes":       # sizes in the local header: the checks that need them
        members[5:5] = [synth.ZipMember("junk.txt", txt, comp=comp + b"JUNKJUNK"),
                        synth.ZipMember("usize.txt", txt, usize=len(txt) + 1),
                        synth.ZipMember("bt3.txt", txt, comp=bytes([comp[0] | 6]) + comp[1:])]
    return synth.make_zip(members, framing=framing)


def many_small(framing):
    """Hundreds of small entries: with sizes in the local headers the streaming reader decodes
    the deflate entries that follow the current one in the same device pass."""
    txt = synth.synth_text(400000, 29)
    members = []
    for i in range(400):
        body = txt[i * 700:i * 700 + 200 + (i * 37) % 3000]
        if i % 9 == 4:
            members.append(synth.ZipMember("s%03d.bin" % i, body, method=0))
        elif i == 77:
            members.append(synth.ZipMember("badcrc%03d" % i, body, crc=12345))
        elif i == 150:
            members.append(synth.ZipMember("empty%03d" % i, b""))
        elif i == 201:
            comp = synth.deflate_raw(body, 6)
            members.append(synth.ZipMember("cut%03d" % i, body, comp=comp[:len(comp) // 2]))
        else:
            members.append(synth.ZipMember("f%03d.txt" % i, body, level=1 + i % 9))
    return synth.make_zip(members, framing=framing)


def report(binary, blob, block, opt=None, raw=False):
    with tempfile.NamedTemporaryFile(suffix=".zip", delete=False) as f:
        f.write(blob)
    try:
        cmd = [binary, "list", f.name, "--meta"] + (["--stream", str(block)] if block else []) + \
              (["--opt", opt] if opt else []) + (["--raw"] if raw else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        return [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    finally:
        os.unlink(f.name)


def comparable(lines):
    """The reference's block sizes in streaming mode follow the input chunks zlib was fed; this
    build decodes an entry whole and serves 256 KiB blocks.  So block counts / sizes are not
    compared, nor - for an entry that FAILS - how many bytes went out before the failing block.
    Names, sizes, modes, times, return codes, messages, and bytes + CRC of good entries are."""
    out = []
    for l in lines:
        d = {k: v for k, v in l.items() if k not in ("nblk", "blocks")}
        if d.get("rd", 1) < 0:
            d.pop("nbytes", None)
            d.pop("crc", None)
        out.append(d)
    return out


def check(ref_binary, new_binary):
    for framing in ("sizes", "at_end"):
        z = archive(framing)
        for block in (977, 65536, 1 << 24):
            for blob, opt in ((z, None), (z, "zip:ignorecrc32"), (z[:len(z) // 3], None)):
                a = comparable(report(ref_binary, blob, block, opt))
                b = comparable(report(new_binary, blob, block, opt))
                assert a == b, (framing, block, opt, len(blob), a, b)
        z = many_small(framing)
        for block in (4096, 1 << 24):
            a = comparable(report(ref_binary, z, block))
            b = comparable(report(new_binary, z, block))
            assert a == b, (framing, block, [x for x, y in zip(a, b) if x != y][:3], [y for x, y in zip(a, b) if x != y][:3])


def check_fixtures(ref_binary, new_binary, fixture_dir, raw_of, refused=()):
    """Every reference fixture, seekable and streamed, with the full metadata report (owner,
    access / change times, link targets, encryption flags): identical on both libraries."""
    for name in sorted(os.listdir(fixture_dir)):
        blob = open(os.path.join(fixture_dir, name), "rb").read()
        raw = bool(raw_of.get(name, False))
        # (an entry this build refuses - encrypted - is only compared in seekable mode: how far the
        # reference's decryption set-up got before failing decides where its streaming reader resumes)
        for block in ((0,) if raw or name in refused else (0, 1000)):
            a = report(ref_binary, blob, block, raw=raw)
            b = report(new_binary, blob, block, raw=raw)
            if block:
                a, b = comparable(a), comparable(b)
            if name in refused:        # refused with this build's own message: same codes, other text
                a = [{k: v for k, v in l.items() if k != "err"} for l in a]
                b = [{k: v for k, v in l.items() if k != "err"} for l in b]
            assert a == b, (name, block, [x for x, y in zip(a, b) if x != y][:2], [y for x, y in zip(a, b) if x != y][:2])


def bgzf_with_lying_trailers(n_members=300, member=60000, seed=7):
    """A BGZF chain whose data is intact but whose ISIZE fields are not: too small (a few
    members, one of them by a lot), too large, zero.  The reference never looks at ISIZE
    (archive_read_support_filter_gzip.c:423), so it decodes all of it."""
    import struct
    parts = synth.split_text(n_members * member, member, seed)
    blob = bytearray(synth.make_bgzf(parts))
    spans, off = [], 0
    while off < len(blob):
        bsize = struct.unpack_from("<H", blob, off + 16)[0] + 1
        spans.append((off, bsize))
        off += bsize

    def set_isize(i, fn):
        o, b = spans[i]
        struct.pack_into("<I", blob, o + b - 4, fn(struct.unpack_from("<I", blob, o + b - 4)[0]) & 0xFFFFFFFF)
    set_isize(0, lambda v: v - 1)            # the first member of the first (small) window
    set_isize(5, lambda v: v - 100)
    set_isize(20, lambda v: v + 77)
    set_isize(41, lambda v: 0)
    set_isize(42, lambda v: 16)
    set_isize(n_members - 60, lambda v: v - 4096)    # in a later, larger window
    set_isize(n_members - 1, lambda v: v // 2)       # the last data member
    return bytes(blob), b"".join(parts)


def check_bgzf_trailers(ref_binary, new_binary, n_members=300):
    """n_members = 300: windows below the size at which b2i_submit stages its input (the filter
    keeps its own copy); 1500: the first window is 8 MiB of input and is found in the staging."""
    blob, plain = bgzf_with_lying_trailers(n_members)
    a = report(ref_binary, blob, 0, raw=True)
    b = report(new_binary, blob, 0, raw=True)
    assert a[0]["rd"] == 1 and a[0]["nbytes"] == len(plain), a
    assert a == b, (a, b)
    # cut in the middle of a member: both stop with the same report
    cut = blob[:len(blob) // 2 + 123]
    a, b = report(ref_binary, cut, 0, raw=True), report(new_binary, cut, 0, raw=True)
    assert comparable(a) == comparable(b), (a, b)
