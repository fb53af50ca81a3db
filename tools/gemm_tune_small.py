"""Tile configurations for the small per-step recurrent GEMMs (M = batch = 512) on the serial critical path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_tune import shape
cf = [6441, 6442, 6462, 6433, 12823, 12842, 12832]
shape("time fwd step", 512, 3072, 1024, "nt", cf)
shape("time bwd dh", 512, 1024, 3072, "nn", cf)
shape("chord dec fwd step", 512, 1536, 512, "nt", cf)
shape("chord dec bwd dh", 512, 512, 1536, "nn", cf)
shape("enc bi-GRU fwd step", 512, 3072, 1024, "nt", cf)
shape("time->notes hid", 16384, 512, 1024, "nt", cf + [25622, 925641])
shape("gi_s", 16384, 1536, 1024, "nt", cf + [25622, 925641])
