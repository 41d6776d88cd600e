"""openlte_b200 -- B200-native LTE downlink receive hot path (drop-in for the liblte_phy functions that
LTE_fdd_dl_file_scan exercises).  The product is the C-ABI library libopenlte_b200.so (include/lte_b200.h)
and its C++ liblte_phy-compatible surface (include/liblte_phy_b200.h); this package only loads it and gives
tests / bench.py typed access.  There is no CPU fallback anywhere in this package."""
from .lib import Lib, load, LibraryMissing, Alloc  # noqa: F401
