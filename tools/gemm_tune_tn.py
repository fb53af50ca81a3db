"""Tile configurations for the HBM-bound weight-gradient (TN, huge K) GEMMs of the training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_tune import shape  # noqa  (runs its own table first when imported as a script; here only the helper)

cf = [12823, 12841, 12832, 6441, 6433, 25622]
shape("summ dW_ih (3,1,99)", 384, 128, 262144, "tn", cf)
shape("summ dW_hh", 384, 128, 262144 - 1, "tn", cf)
shape("note dW_ih tok", 1536, 128, 245760, "tn", cf)
shape("note dW_hh", 1536, 512, 245760, "tn", cf)
shape("time dW_hh (24,8,1)", 3072, 1024, 16384, "tn", cf)
shape("time dW_ih", 3072, 256, 16384, "tn", cf)
shape("pitch head dW", 136, 512, 245760, "tn", cf)
shape("dur GX^T S", 264, 72, 1474560, "tn", cf)
