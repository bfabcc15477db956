#!/usr/bin/env python3
"""Generate tests/golden/ from the reference's own hot-path fixtures.

Run in the build container (needs /root/reference and oracle/_ref built by
`make -C oracle`).  For every ZIP / gzip fixture the reference's tests use on
this path it

  1. uudecodes `<reference>/libarchive/test/<name>.uu` into
     tests/golden/ref_fixtures/<name>   (small binary test vectors, not source),
  2. runs oracle/_ref/oracle_extract (the UNMODIFIED reference + system zlib)
     on it and stores the per-entry report (names, sizes, return codes, CRC-32
     of the bytes read, block sizes, error strings) in
     tests/golden/ref_expected.json, plus sha256 of the concatenated data.

The GPU box has no /root/reference: tests read only the committed outputs.
"""
import binascii
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST = "/root/reference/libarchive/test"
OUT = os.path.join(ROOT, "tests", "golden")
EXTRACT = os.path.join(ROOT, "oracle", "_ref", "oracle_extract")

ZIP_FIXTURES = [
    "test_read_format_zip.zip",
    "test_read_format_zip_ux.zip",
    "test_read_format_zip_length_at_end.zip",
    "test_read_format_zip_7z_deflate.zip",
    "test_read_format_zip_high_compression.zip",
    "test_read_format_zip_zip64a.zip",
    "test_read_format_zip_zip64b.zip",
    "test_read_format_zip_padded1.zip",
    "test_read_format_zip_padded2.zip",
    "test_read_format_zip_padded3.zip",
    "test_read_format_zip_sfx",
    "test_read_format_zip_extra_padding.zip",
    "test_read_format_zip_malformed1.zip",
    "test_read_format_zip_nested.zip",
    "test_read_format_zip_with_invalid_traditional_eocd.zip",
    "test_read_format_zip_comment_stored_1.zip",
    "test_read_format_zip_comment_stored_2.zip",
    "test_read_format_zip_msdos.zip",
    "test_read_format_zip_nofiletype.zip",
    "test_read_format_zip_symlink.zip",
    "test_read_format_zip_filename_utf8_ru.zip",
    "test_compat_zip_1.zip",
    "test_compat_zip_2.zip",
    "test_compat_zip_3.zip",
    "test_compat_zip_4.zip",
    "test_compat_zip_5.zip",
    "test_compat_zip_6.zip",
    "test_compat_zip_7.xps",
    "test_compat_zip_8.zip",
    # detection of what this build must refuse (encrypted / non-deflate)
    "test_read_format_zip_encryption_data.zip",
    "test_read_format_zip_bzip2.zipx",
]
GZIP_FIXTURES = [
    "test_compat_gzip_1.tgz",
    "test_compat_gzip_2.tgz",
    "test_read_format_raw.data.gz",
]


def uudecode(path):
    out = bytearray()
    started = False
    with open(path, "rb") as f:
        for line in f:
            if not started:
                if line.startswith(b"begin "):
                    started = True
                continue
            s = line.rstrip(b"\r\n")
            if s == b"end":
                break
            if not s or s == b"`":
                continue
            try:
                out += binascii.a2b_uu(s)
            except binascii.Error:
                n = (((s[0] - 32) & 63) * 4 + 5) // 3
                out += binascii.a2b_uu(s[:n])
    return bytes(out)


def run_extract(path, raw):
    dump = path + ".dump"
    cmd = [EXTRACT, "list", path, "--dump", dump]
    if raw:
        cmd.append("--raw")
    r = subprocess.run(cmd, capture_output=True, text=True, check=True, timeout=20)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.strip()]
    with open(dump, "rb") as f:
        data = f.read()
    os.unlink(dump)
    return lines, hashlib.sha256(data).hexdigest(), len(data)


def main():
    fx = os.path.join(OUT, "ref_fixtures")
    os.makedirs(fx, exist_ok=True)
    expected = {}
    for name, raw in [(n, False) for n in ZIP_FIXTURES] + [(n, True) for n in GZIP_FIXTURES]:
        src = os.path.join(REF_TEST, name + ".uu")
        if not os.path.exists(src):
            print("missing", src, file=sys.stderr)
            continue
        blob = uudecode(src)
        dst = os.path.join(fx, name)
        with open(dst, "wb") as f:
            f.write(blob)
        lines, sha, n = run_extract(dst, raw)
        expected[name] = {"raw": raw, "file_sha256": hashlib.sha256(blob).hexdigest(),
                          "data_sha256": sha, "data_bytes": n, "report": lines}
        print(f"{name}: {len(blob)} B, {len(lines) - 1} entries, {n} data bytes")
    with open(os.path.join(OUT, "ref_expected.json"), "w") as f:
        json.dump(expected, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
