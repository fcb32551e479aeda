"""B200-native SAM 2.1 mask-propagation hot path (memory attention, mask decoder, memory encoder,
connected components) behind the reference's `sam2` module/predictor API.  All arithmetic on the
path runs in hand-written sm_100a CUDA inside `libvls_b200.so` (C ABI in include/vls_b200.h);
this package is the Python host side that mirrors the reference interfaces."""
__version__ = "0.1.0"
