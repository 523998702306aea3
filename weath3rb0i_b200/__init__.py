"""weath3rb0i_b200 — B200-native block-parallel codec for weath3rb0i's
predict→update→code path (reference: src/main.rs:89-152).

Host-side mirror of the reference's interface for this path:
`compress(input_file, output_file)`, `decompress(input_file, output_file)` and
`run(path, action)` (src/main.rs:55-144), on top of the C ABI in
include/w3_codec.h.  All coding happens in CUDA kernels; there is no CPU path.
"""
from .capi import Context, MagicMismatch, W3Error, DEFAULT_BLOCK  # noqa: F401
from .codec import Action, compress, decompress, run, default_context  # noqa: F401
