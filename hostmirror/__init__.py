"""hostmirror — the reference crate's OWN Rust around the C-ABI boundary, restated in Python because this image has no Rust toolchain.

Not the product: the product boundary is `libwdr_b200.{so,a}` + `include/wdr.h` (+ the ctypes binding `whisper-diarize-rs_b200/capi.py`), and in
a deployment the crate's Rust stays the host.  These modules exist so that tests and bench.py can drive the library the way
the crate does and check the consumer contract:

* host.py       — src/vad.rs:33-82 (mask / merge / slice), src/transcribe.rs:171-320 (token -> word timestamps), :397-459 (offsets,
                  overlap clipping), :461-497 (speaker policy), src/utils.rs:3-59; function for function, same names.
* formatting.py — src/formatting.rs `process_segments` (the CONSUMER of the library's output), pinned by the
                  reference's own `segments.json` fixture (tests/test_formatting.py).
"""
from . import host, formatting  # noqa: F401
