"""Graph-replayed training step: note-summary bi-GRU on the weight-resident kernel vs the per-step tcgen05 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from chunk_sweep import timed, ops  # noqa  (chunk_sweep runs its own table on import)

for rows in (4096, 1 << 30, 4096, 1 << 30):
    ops.RESIDENT_GRU128_MAX_ROWS_TF32 = rows
    timed(f"resident GRU128 up to {rows} sequences")
