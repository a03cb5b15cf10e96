"""Drop-in for turlando/flac-py's `flac` package whose encode / decode hot path runs on a
B200 through libflacb200.so (hand-written sm_100a CUDA kernels behind a ctypes C ABI).

Same module names and entry points as the reference (flac/encoder.py:48 `encode`,
flac/encoder.py:33 `EncoderParameters`, flac/decoder.py:31 `decode`, `python -m flac`),
there is no CPU fallback: importing works anywhere, encoding or decoding without the CUDA
library and a GPU raises.
"""
__all__ = ["encoder", "decoder", "binary", "crc", "coded_number", "utils", "common"]
