"""ltesniffer_b200 -- B200-native (sm_100a) LTE PHY decode path behind a C-ABI.

The product is libltephy_b200.so (hand-written CUDA kernels + C++ host code, see csrc/ and
include/ltephy_b200.h).  This package only builds and binds it; it contains no CPU implementation of
the path and raises if the library cannot be loaded."""
from .capi import LtePhy, Search, decode_subframes, load_library, Grant, TbResult, Cand, SfInfo, Cfg  # noqa: F401
