"""prrn_aln_b200 -- B200-native (sm_100a) dynamic-programming forward fill of ogotoh/prrn_aln.

The product is libprrn_gpu.so (hand-written CUDA behind the C ABI of include/prrn_gpu.h); this
package is the thin host-side mirror of the reference's call surface for that path
(alnScoreD / calcdist / ...) used by the tests, bench.py and multi-GPU sharding.  There is no CPU
fallback: calls raise PgError when the CUDA library or a device is missing.
"""
from .api import (ALPRM, Context, PgError, Params, SeqSet, alnScoreD, calcdist, calcdist_cells,  # noqa: F401
                  elem, lib_path, load_library, declared_symbols, align2, stdskl, packed_plan_coverage,
                  GParams, group_cells, gparams_from_pwd, alignB_ng)
from . import groups  # noqa: F401
