"""libarchive_b200 — B200-native (sm_100a) inflate + CRC-32 for libarchive's
ZIP / BGZF read path.  The product is libb200inflate.so (C ABI in
include/b200inflate.h) and the libarchive plugins under csrc/plugin/; this
package is the Python harness around that ABI for tests and benchmarks."""
from . import capi, reader, synth  # noqa: F401

__all__ = ["capi", "reader", "synth"]
