"""Mutated ZIP / BGZF inputs against the AddressSanitizer build of the plugins' host logic
(make -C tests/refsuite asan): one-call, streaming-engine and file-backed modes, seekable and
streamed.  No GPU: the C ABI is answered by the oracle shim.  usage: asan_fuzz_hostlogic.py [seed] [inputs]"""
import os, random, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
EXE = os.path.join(ROOT, 'tests', 'refsuite', '_out', 'asan', 'extract_asan')
from libarchive_b200 import synth
import stream_cases
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bases = [stream_cases.archive("sizes"), stream_cases.archive("at_end"), stream_cases.many_small("sizes")[:200000],
         synth.config1_zip_text64k(40, 20000), synth.make_bgzf(synth.split_text(30 * 30000, 30000, 3))]
envs = [{}, {"B2I_ZIP_PIPE": "1", "B2I_PIPE_WINDOW_MB": "1"}, {"B2I_ZIP_PIPE": "1", "B2I_ZIP_FILE_MODE": "1", "B2I_PIPE_WINDOW_MB": "1"}]
bad = 0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for it in range(N):
    bi = random.randrange(len(bases))
    b = bytearray(bases[bi])
    mode = random.random()
    if mode < 0.7:
        for _ in range(random.randint(1, 8)):
            # bias towards the tail (directory) and the head (first local header)
            r = random.random()
            pos = random.randrange(len(b)) if r < 0.4 else (len(b) - 1 - random.randrange(min(len(b), 2000)) if r < 0.8 else random.randrange(min(len(b), 200)))
            b[pos] = random.randrange(256)
    elif mode < 0.85:
        b = b[:random.randrange(1, len(b))]
    else:
        p = random.randrange(len(b)); b[p:p] = bytes(random.randrange(256) for _ in range(random.randint(1, 64)))
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(b)
    try:
        for env in envs:
            for extra in ([], ["--stream", "977"]):
                cmd = [EXE, "list", f.name] + extra + (["--raw"] if bi == 4 else [])
                r = subprocess.run(cmd, capture_output=True, timeout=120,
                                   env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=0", **env))
                if r.returncode not in (0, 1, 2) or b"AddressSanitizer" in r.stderr:
                    bad += 1
                    keep = "/tmp/asan_crash_%d_%d.bin" % (it, bad)
                    open(keep, "wb").write(b)
                    print("CRASH", it, env, extra, r.returncode, keep, r.stderr[-1500:].decode(errors="replace"), flush=True)
    finally:
        os.unlink(f.name)
print("done", N, "inputs, bad:", bad)
