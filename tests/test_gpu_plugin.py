"""The drop-in itself: libarchive_dropin.so = the reference's own read core and every
other object, compiled unmodified, + the two B200 plugin modules (csrc/plugin/).
The SAME driver source (oracle/oracle_extract.c, public libarchive API only) is linked
once against the unmodified reference (oracle/_ref) and once against the drop-in; their
per-entry reports must be identical: names, sizes, modes, mtimes, return codes, error
strings, block sizes, CRC of the bytes read, and the bytes themselves."""
import hashlib
import json
import os
import subprocess
import tempfile
import zlib

import pytest

import oracle_binding as ob
from libarchive_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "libarchive_b200", "dropin_extract")
GOLD = os.path.join(ROOT, "tests", "golden")
with open(os.path.join(GOLD, "ref_expected.json")) as f:
    EXPECTED = json.load(f)

# what this build refuses
REFUSED = {"test_read_format_zip_encryption_data.zip"}       # different (still FAILED) message


def run(binary, path, raw=False, opt=None, env=None, stream=0, file=False):
    dump = path + ".dump." + os.path.basename(binary)
    cmd = [binary, "list", path, "--dump", dump]
    if file:
        cmd.append("--file")                  # archive_read_open_filename instead of open_memory
    if stream:
        cmd += ["--stream", str(stream)]      # read callback only: the ZIP streaming reader runs
    if raw:
        cmd.append("--raw")
    if opt:
        cmd += ["--opt", opt]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120,
                       env=dict(os.environ, **env) if env else None)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    data = b""
    if os.path.exists(dump):
        with open(dump, "rb") as f:
            data = f.read()
        os.unlink(dump)
    return lines, data


def need_dropin():
    if not os.path.exists(DROPIN):
        pytest.skip("drop-in libarchive not built (needs /root/reference at build time)")


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_reference_fixture_through_the_dropin(name):
    need_dropin()
    exp = EXPECTED[name]
    path = os.path.join(GOLD, "ref_fixtures", name)
    lines, data = run(DROPIN, path, raw=exp["raw"])
    want = exp["report"]
    assert len(lines) == len(want), (lines, want)
    for g, w in zip(lines, want):
        if name in REFUSED and "err" in w and w["rd"] == -25:
            assert g["rd"] == -25
            g = dict(g, err=w["err"])
        assert g == w, (name, g, w)
    if name not in REFUSED:
        assert hashlib.sha256(data).hexdigest() == exp["data_sha256"]


def both(blob, raw=False, opt=None, env=None, stream=0, file=False):
    need_dropin()
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(blob)
    try:
        a = run(ob.REF_EXTRACT, f.name, raw, opt, None, stream, file)
        b = run(DROPIN, f.name, raw, opt, env, stream, file)
    finally:
        os.unlink(f.name)
    return a, b


def _mixed_members():
    parts = synth.split_text(200 * 65536, 65536, 15)
    members = [synth.ZipMember("dir/e%04d.txt" % i, p, level=1 + i % 9) for i, p in enumerate(parts)]
    big = synth.synth_text(3 << 20, 16)
    members += [synth.ZipMember("big.txt", big), synth.ZipMember("fixed", big[:400000], strategy=zlib.Z_FIXED),
                synth.ZipMember("stored.bin", synth.synth_random(700001, 1), method=0),
                synth.ZipMember("bad-crc.txt", big[:70000], crc=0x1234),
                synth.ZipMember("short-usize", big[:9000], usize=5000),
                synth.ZipMember("rnd.def", synth.synth_random(300000, 2)), synth.ZipMember("empty", b""),
                synth.ZipMember("d/", b"", method=0)]
    members += [synth.ZipMember("tail/e%04d.txt" % i, p[:30000]) for i, p in enumerate(parts[:40])]
    return members


@pytest.mark.parametrize("mode", ["pipe-memory", "file-windows", "file-small-image"])
def test_streaming_engine_identical_reports(mode):
    """The plugin's streaming mode (b2i_pipe: windows through the pinned ring) and its
    file-backed mode (archive_read_open_filename read window by window through
    __archive_read_seek / _ahead): same reports and bytes as the reference, including the
    error entries, an overflowing entry and stored entries served from the staged input."""
    members = _mixed_members()
    blob = synth.make_zip(members, framing="at_end" if mode == "file-windows" else "sizes")
    env = {"B2I_PIPE_WINDOW_MB": "2"}
    if mode == "pipe-memory":
        env["B2I_ZIP_PIPE"] = "1"
        (ra, da), (rb, db) = both(blob, env=env)
    elif mode == "file-windows":
        env["B2I_ZIP_FILE_MODE"] = "1"
        (ra, da), (rb, db) = both(blob, env=env, file=True)
    else:
        (ra, da), (rb, db) = both(blob, file=True)
    if mode == "pipe-memory":
        assert ra == rb and da == db
        return
    # A file source reaches the reference in 64 KiB blocks, and its zlib loop ends a block
    # wherever the input block ends: block boundaries (and how much of a FAILING entry was
    # handed out before its end-of-entry check) follow the I/O, not the archive.  Everything
    # else - names, sizes, return codes, messages, the CRC of every delivered body - is equal.
    def norm(r):
        r = {k: v for k, v in r.items() if k not in ("nblk", "blocks")}
        if r.get("rd") != 1:
            r.pop("nbytes", None), r.pop("crc", None)
        return r
    assert [norm(r) for r in ra] == [norm(r) for r in rb]


@pytest.mark.parametrize("framing", ["sizes", "at_end"])
def test_generated_zip_identical_reports(framing):
    parts = synth.split_text(300 * 65536, 65536, 5)
    members = [synth.ZipMember("dir/e%04d.txt" % i, p, level=1 + i % 9) for i, p in enumerate(parts)]
    big = synth.synth_text(3 << 20, 6)
    members += [synth.ZipMember("big.txt", big), synth.ZipMember("fixed", big[:400000], strategy=zlib.Z_FIXED),
                synth.ZipMember("stored.bin", synth.synth_random(700001, 1), method=0),
                synth.ZipMember("rnd.def", synth.synth_random(300000, 2)), synth.ZipMember("empty", b""),
                synth.ZipMember("d/", b"", method=0)]
    (ra, da), (rb, db) = both(synth.make_zip(members, framing=framing))
    assert ra == rb
    assert da == db and hashlib.sha256(da).hexdigest() == hashlib.sha256(b"".join(m.data for m in members)).hexdigest()


def test_error_entries_identical_reports():
    txt = synth.synth_text(600000, 3)
    comp = synth.deflate_raw(txt, 6)
    crc = zlib.crc32(txt) & 0xFFFFFFFF
    members = [synth.ZipMember("ok", txt), synth.ZipMember("cut", txt, comp=comp[:len(comp) // 2]),
               synth.ZipMember("bt3", txt, comp=bytes([comp[0] | 6]) + comp[1:]),
               synth.ZipMember("midflip", txt, comp=comp[:100000] + bytes([comp[100000] ^ 0x10]) + comp[100001:]),
               synth.ZipMember("junk", txt, comp=comp + b"JUNKJUNK"), synth.ZipMember("crc", txt, crc=crc ^ 1),
               synth.ZipMember("usize", txt, usize=len(txt) + 1), synth.ZipMember("after", txt[:1000])]
    (ra, da), (rb, db) = both(synth.make_zip(members))
    assert ra == rb and da == db
    (ra, da), (rb, db) = both(synth.make_zip(members), opt="zip:ignorecrc32")
    assert ra == rb and da == db


def test_bgzf_and_plain_gzip_identical_reports():
    parts = synth.split_text(500 * 65280, 65280, 9)
    (ra, da), (rb, db) = both(synth.make_bgzf(parts) + b"trailing garbage", raw=True)
    assert ra == rb and da == db == b"".join(parts)
    a, b = synth.synth_text(200000, 1), synth.synth_random(5000, 2)
    f = synth.gzip_member(a, name=b"a.txt", mtime=1234) + synth.gzip_member(b, comment=b"c", hcrc=True) + \
        synth.gzip_member(b"", extra=b"XY\x02\x00zz")
    (ra, da), (rb, db) = both(f, raw=True)
    assert ra == rb and da == db == a + b
    (ra, da), (rb, db) = both(f[:len(f) // 3], raw=True)          # truncated inside member 1
    assert ra == rb and da == db
    bad = bytearray(f); bad[len(f) // 4] ^= 0x40                      # corrupt member 1's deflate data
    # like the reference (gzip.c:423) the drop-in does not check the gzip trailer by default:
    # identical reports; B2I_GZIP_VERIFY=1 makes it notice the CRC / ISIZE mismatch
    (ra, da), (rb, db) = both(bytes(bad), raw=True)
    assert ra == rb and da == db
    (ra, da), (rb, db) = both(bytes(bad), raw=True, env={"B2I_GZIP_VERIFY": "1"})
    if ra[0].get("rd") == 1:
        assert rb[0].get("rd", rb[0].get("open")) == -30


def test_bgzf_members_with_wrong_isize_identical_reports():
    """ISIZE too small / too large / zero in some BGZF trailers (data intact): the reference
    ignores the field; the drop-in decodes such a member again with room."""
    need_dropin()
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    import stream_cases
    stream_cases.check_bgzf_trailers(ob.REF_EXTRACT, DROPIN, 300)
    stream_cases.check_bgzf_trailers(ob.REF_EXTRACT, DROPIN, 1500)


def test_streaming_reader_identical_reports():
    """Non-seekable input: the streaming ZIP reader (local headers, data descriptors) of the
    drop-in against the reference's, on generated archives with good and bad entries."""
    need_dropin()
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    import stream_cases
    stream_cases.check(ob.REF_EXTRACT, DROPIN)


def test_reference_fixtures_full_metadata_identical():
    """All reference fixtures through the drop-in and the reference, seekable and streamed, with
    owner, access / change times, link targets and encryption flags in the report."""
    need_dropin()
    if not ob.have_ref():
        pytest.skip("oracle/_ref not built")
    import stream_cases
    stream_cases.check_fixtures(ob.REF_EXTRACT, DROPIN, os.path.join(GOLD, "ref_fixtures"),
                                {k: v["raw"] for k, v in EXPECTED.items()}, refused=tuple(REFUSED))


def test_handles_on_several_threads(tmp_path):
    """Six threads, each with its own archive handle, read the same archive through the drop-in
    at the same time (tests/csrc/mt_read.c): one-call path or streaming engine as the size
    decides, the engine forced with small windows, and the BGZF filter.  Every thread sees the
    same bytes (context pool, pinned shelf, copy threads, five jobs per context under
    concurrency)."""
    exe = os.path.join(ROOT, "tests", "refsuite", "_out", "mt_read_dropin")
    if not os.path.exists(exe):
        pytest.skip("tests/refsuite/_out/mt_read_dropin not built (needs /root/reference)")
    parts = synth.split_text(600 * 50000, 50000, 31)
    want = "%08x" % (zlib.crc32(b"".join(parts)) & 0xFFFFFFFF)
    z = tmp_path / "a.zip"
    z.write_bytes(synth.make_zip([synth.ZipMember("e%03d" % i, p) for i, p in enumerate(parts)]))
    g = tmp_path / "a.bgzf"
    g.write_bytes(synth.make_bgzf(parts))
    for path, extra, env in ((z, [], {}), (z, [], {"B2I_ZIP_PIPE": "1", "B2I_PIPE_WINDOW_MB": "4"}), (g, ["--raw"], {})):
        r = subprocess.run([exe, str(path), "6", "3"] + extra, capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, (env, r.stdout, r.stderr[-500:])
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert j["bad_threads"] == 0 and j["bytes"] == 600 * 50000 and j["crc"] == want, (env, j)
