"""Run the reference's own libarchive_test programs (hot-path subset, built by
tests/refsuite/Makefile from the unmodified reference sources) against the
unmodified reference library and against the drop-in whose ZIP format and gzip
filter modules run on the GPU; one process per test so that a crash is one
result.  Prints / returns {test: (ref_status, dropin_status)}.

Test infrastructure: compares, never ships."""
from __future__ import annotations

import json
import os
import time
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFDIR = os.path.join(ROOT, "oracle", "_ref")               # built from the reference alone
NEWDIR = os.path.join(ROOT, "tests", "refsuite", "_out")     # linked against this repo's modules
DATADIR = os.path.join(ROOT, "tests", "refsuite", "testdata")  # the reference's .uu fixtures these tests name


def list_tests(binary):
    out = subprocess.run([binary, "-l"], capture_output=True, text=True, timeout=60).stdout
    # "  0: test_name"
    return [m.group(1) for m in re.finditer(r"^\s*\d+:\s+(\S+)", out, re.M)]


def run_one(binary, test, timeout=300):
    with tempfile.TemporaryDirectory() as tmp:
        try:
            p = subprocess.run([binary, "-r", DATADIR, "-q", test], cwd=tmp,
                               capture_output=True, text=True, timeout=timeout,
                               env=dict(os.environ, TMPDIR=tmp))
        except subprocess.TimeoutExpired:
            return "timeout", ""
        text = p.stdout + p.stderr
        for dp, _, fns in os.walk(tmp):       # the runner writes assertion details to <test>.log
            for fn in fns:
                if fn.endswith(".log"):
                    text += open(os.path.join(dp, fn), errors="replace").read()[:6000]
        if p.returncode < 0:
            return f"signal{-p.returncode}", text[-8000:]
        failed = re.search(r"Tests failed:\s+(\d+)", text)
        skipped = re.search(r"Skips reported:\s+(\d+)", text)
        asserts = re.search(r"Assertions failed:\s+(\d+)", text)
        if p.returncode == 0 and (failed is None or failed.group(1) == "0"):
            return ("skipped" if "skipped" in text.split("Totals:")[0].lower() and skipped and skipped.group(1) != "0"
                    and re.search(r"Assertions checked:\s+0\b", text) else "ok"), ""
        return f"FAIL({asserts.group(1) if asserts else '?'})", text[-8000:]


# Tests the drop-in is NOT expected to pass, with the reason: features
# outside the hot path that this build refuses instead of decoding on the CPU.
KNOWN_GAPS = {
    "test_read_format_zip_mac_metadata": "Mac resource-fork folding (mac-ext) not provided",
    "test_read_format_zip_ppmd8_crash_1": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_ppmd8_crash_2": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_ppmd_multi": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_ppmd_multi_blockread": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_ppmd_one_file": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_ppmd_one_file_blockread": "ZIPX PPMd (method 98) not provided",
    "test_read_format_zip_traditional_encryption_data": "PKWARE decryption not provided",
    "test_write_format_zip_traditional_pkware_encryption": "PKWARE decryption not provided",
}


def main(argv, variant="dropin"):
    ref = os.path.join(REFDIR, "libarchive_test_ref")
    drop = os.path.join(NEWDIR, "libarchive_test_" + variant)
    only = [x for x in argv[1:] if not x.startswith("-")]
    skip = [x[1:] for x in argv[1:] if x.startswith("-")]
    tests = [t for t in list_tests(ref) if (not only or any(o in t for o in only)) and not any(k in t for k in skip)]
    results, logs = {}, {}
    for t in tests:
        t0 = time.time()
        r, _ = run_one(ref, t)
        t1 = time.time()
        d, log = run_one(drop, t)
        t2 = time.time()
        results[t] = (r, d, round(t1 - t0, 2), round(t2 - t1, 2))
        if d != r:
            logs[t] = log
        gap = "  (declared gap: %s)" % KNOWN_GAPS[t] if t in KNOWN_GAPS and d != r else ""
        print(f"{t:60s} ref={r:8s} {t1 - t0:6.1f}s   {variant}={d:10s} {t2 - t1:6.1f}s{gap}", flush=True)
    same = sum(1 for v in results.values() if v[0] == v[1])
    print(f"== {same} of {len(results)} tests give the same verdict on both libraries")
    return results, logs


if __name__ == "__main__":
    variant = "dropin"
    if "--hostlogic" in sys.argv:
        sys.argv.remove("--hostlogic")
        variant = "hostlogic"
    res, logs = main(sys.argv, variant)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump({"results": res, "logs": logs}, open(os.path.join(out, "refsuite.json"), "w"), indent=1)
