"""B200-native DR-SPAAM per-point scan hot path (cutout -> attention memory -> centre NMS).

Host side mirrors the reference's Python API; the arithmetic runs in libpof.so
(hand-written sm_100a CUDA behind the C ABI of include/pof.h).  No CPU fallback.
"""
__version__ = "0.1.0"
