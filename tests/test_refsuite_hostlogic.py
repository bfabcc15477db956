"""The reference's OWN test programs (libarchive_test, hot-path subset: every test that reads
ZIP or gzip data, plus the write-then-read tests) run against the plugin modules' HOST logic.

tests/refsuite/Makefile compiles the reference's test sources unmodified and links them with
the reference's other objects + this repo's three plugin modules + tests/emul/b2i_shim.c, which
answers the C ABI with the CPU oracle (test infrastructure only; the product has no such path).
What is exercised here is everything around the device call: header walking in both readers
(seekable and streaming), data descriptors, the block contract, error mapping, options, the
gzip filter's member handling.  The same programs run against the real drop-in on the GPU in
test_gpu_refsuite.py."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "refsuite"))
from run_refsuite import DATADIR, KNOWN_GAPS, NEWDIR, REFDIR  # noqa: E402


def run_all(binary, tmp_path, timeout):
    p = subprocess.run([binary, "-q", "-r", DATADIR], cwd=tmp_path, capture_output=True,
                       text=True, errors="replace", timeout=timeout, env=dict(os.environ, TMPDIR=str(tmp_path)))
    text = p.stdout + p.stderr
    failing = set(re.findall(r"^\s*\d+: (\S+) \(\d+ failures\)", text, re.M))
    # quiet mode prints one '.' or 'E' per test after "Running "
    prog = re.search(r"^Running ([.E\n]+)", text, re.M)
    n = sum(prog.group(1).count(c) for c in ".E") if prog else 0
    return p.returncode, n, failing, text


def test_reference_tests_pass_on_plugin_host_logic(tmp_path):
    binary = os.path.join(NEWDIR, "libarchive_test_hostlogic")
    if not os.path.exists(binary):
        pytest.skip("tests/refsuite/_out/libarchive_test_hostlogic not built (needs /root/reference)")
    rc, n, failing, text = run_all(binary, tmp_path, 900)
    assert rc >= 0, "the test runner crashed:\n" + text[-2000:]
    assert n >= 125, text[-2000:]
    unexpected = failing - set(KNOWN_GAPS)
    assert not unexpected, "reference tests failing outside the declared gaps: %s" % sorted(unexpected)


def test_streaming_reader_reports_match_reference():
    """Generated archives read through a read callback only (no seeking): the streaming reader's
    host logic against the unmodified reference - good entries, data descriptors, bad CRC, wrong
    sizes, junk after the stream, invalid block type, truncation."""
    ref, new = os.path.join(REFDIR, "oracle_extract"), os.path.join(NEWDIR, "hostlogic_extract")
    if not (os.path.exists(ref) and os.path.exists(new)):
        pytest.skip("oracle/_ref drivers not built (needs /root/reference)")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stream_cases
    stream_cases.check(ref, new)


def test_bgzf_members_with_wrong_isize_match_reference():
    """ISIZE too small / too large / zero in some BGZF trailers: the reference ignores the field,
    the filter decodes such a member again with room (both ways of finding its input again)."""
    ref, new = os.path.join(REFDIR, "oracle_extract"), os.path.join(NEWDIR, "hostlogic_extract")
    if not (os.path.exists(ref) and os.path.exists(new)):
        pytest.skip("oracle/_ref drivers not built (needs /root/reference)")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stream_cases
    stream_cases.check_bgzf_trailers(ref, new, 300)
    stream_cases.check_bgzf_trailers(ref, new, 1500)


def test_reference_fixtures_full_metadata_match_reference():
    """All reference fixtures, seekable and streamed, including owner / access and change times /
    link targets / encryption flags (the extra fields 0x5455, 0x5855, 0x7855, 0x7875, 0x7075)."""
    import json
    ref, new = os.path.join(REFDIR, "oracle_extract"), os.path.join(NEWDIR, "hostlogic_extract")
    if not (os.path.exists(ref) and os.path.exists(new)):
        pytest.skip("oracle/_ref drivers not built (needs /root/reference)")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stream_cases
    gold = os.path.join(ROOT, "tests", "golden")
    expected = json.load(open(os.path.join(gold, "ref_expected.json")))
    stream_cases.check_fixtures(ref, new, os.path.join(gold, "ref_fixtures"), {k: v["raw"] for k, v in expected.items()},
                                refused=("test_read_format_zip_encryption_data.zip",))


def test_handles_on_several_threads(tmp_path):
    """Six threads, each with its own archive handle, read the same archive at the same time
    (one-call path, streaming engine, two "devices", BGZF filter): every thread sees the same
    bytes.  The ThreadSanitizer build of the same program (make -C tests/refsuite tsan) reports
    no race in this repo's code."""
    exe = os.path.join(NEWDIR, "mt_read_hostlogic")
    if not os.path.exists(exe):
        pytest.skip("tests/refsuite/_out/mt_read_hostlogic not built (needs /root/reference)")
    sys.path.insert(0, ROOT)
    import json
    import zlib
    from libarchive_b200 import synth
    parts = synth.split_text(200 * 50000, 50000, 31)
    z = tmp_path / "a.zip"
    z.write_bytes(synth.make_zip([synth.ZipMember("e%03d" % i, p) for i, p in enumerate(parts)]))
    g = tmp_path / "a.bgzf"
    g.write_bytes(synth.make_bgzf(parts))
    want = "%08x" % (zlib.crc32(b"".join(parts)) & 0xFFFFFFFF)
    for path, extra, env in ((z, [], {}), (z, [], {"B2I_ZIP_PIPE": "1", "B2I_PIPE_WINDOW_MB": "1"}),
                             (z, [], {"B2I_ZIP_PIPE": "1", "B2I_PIPE_WINDOW_MB": "1", "B2I_SHIM_GPUS": "2", "B2I_PLUGIN_GPUS": "2"}),
                             (g, ["--raw"], {})):
        r = subprocess.run([exe, str(path), "6", "2"] + extra, capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, (env, r.stdout, r.stderr[-500:])
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert j["bad_threads"] == 0 and j["bytes"] == 200 * 50000 and j["crc"] == want, (env, j)
