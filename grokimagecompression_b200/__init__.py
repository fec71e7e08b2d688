"""grokimagecompression_b200 -- Grok's JPEG 2000 tile-coding hot path (DC shift, RCT/ICT, 5/3 and 9/7
DWT, quantisation, EBCOT Tier-1) as hand-written sm_100a CUDA kernels behind a C ABI.

The product is `libgrok_b200.so` (csrc/, include/grok_b200.h); this package is the thin host-side
mirror used by the tests and bench.py.  There is no CPU fallback.
"""
from .binding import (CBLK_DEC_DTYPE, CBLK_ENC_DTYPE, CBLK_INFO_DTYPE, CBLK_SEG_DTYPE, T1_BLOCK_DTYPE, CompParams, Context,
                      GrokB200Error, Plan, TileParams, lib, LIB_PATH, SYMBOLS, ABI_VERSION)
from . import params

__all__ = ["Context", "Plan", "CompParams", "TileParams", "GrokB200Error", "lib", "params", "LIB_PATH", "SYMBOLS", "ABI_VERSION",
           "CBLK_ENC_DTYPE", "CBLK_DEC_DTYPE", "CBLK_INFO_DTYPE", "CBLK_SEG_DTYPE", "T1_BLOCK_DTYPE"]
