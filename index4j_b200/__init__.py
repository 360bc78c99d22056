"""index4j_b200 — B200-native batched query engine for index4j FM-indexes.

Only the hot path of dynatrace-oss/index4j is here: batched ``FmIndex`` count / locate / extract /
extractUntilBoundary over a serialized index, executed by hand-written sm_100a CUDA kernels behind
the C ABI declared in ``include/fmgpu.h``.  ``FmIndex`` in :mod:`index4j_b200.fm_index` mirrors the
method set of the reference's ``com.dynatrace.fm.FmIndex``
(indices/src/main/java/com/dynatrace/fm/FmIndex.java:443-983).
"""

from .builder import FmIndexBuilder, build_index, gen_log_text, gen_patterns  # noqa: F401
from .fm_index import FmIndex, FmIndexError  # noqa: F401
from .structures import RrrVector, WaveletFixedBlockBoosting  # noqa: F401

__all__ = ["FmIndex", "FmIndexError", "FmIndexBuilder", "build_index", "gen_log_text", "gen_patterns", "RrrVector",
           "WaveletFixedBlockBoosting"]
