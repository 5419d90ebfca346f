"""nanoranger_b200 -- B200-native cell-barcode matcher + UMI collapse for nanoranger.

Drop-in for the stage of mehdiborji/nanoranger that aligns barcode/UMI candidates against the
N-padded whitelist with STAR (scripts/barcode_ref.sh, scripts/barcode_align.sh) and parses the
result (utils.process_matching_*).  Host layer in Python with the reference's function
signatures (nanoranger_b200.utils), compute in hand-written sm_100a kernels behind a C ABI
(include/nanoranger_b200.h, nanoranger_b200/csrc).
"""
from ._lib import (NR_FLAG_BELOW, NR_FLAG_EXHAUSTIVE, NR_FLAG_NO_UMI, NR_FLAG_RC, NR_FLAG_TIE,
                   NR_FLAG_TOO_LONG, NR_MODE_AUTO, NR_MODE_EXHAUSTIVE, NR_MODE_FILTERED,
                   NR_SCORE_BELOW, NR_UMI_NONE)
from .matcher import MatchResult, Whitelist, int_peak, pack_ascii  # noqa: F401

__version__ = "0.1.0"
