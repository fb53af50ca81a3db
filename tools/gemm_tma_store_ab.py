"""A/B: persistent GEMM with TMA bulk-store epilogue (cfg 925641) vs the plain st.global epilogue (dbg bit 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_tune import shape
cf = [925641, 4925641]
shape("gi_tok (output-bound)", 262144, 1536, 128, "nt", cf)
shape("emb gi (output-bound)", 262144, 384, 128, "nt", cf)
shape("note fwd step", 16384, 1536, 512, "nt", cf)
shape("pitch head", 245760, 136, 512, "nt", cf)
shape("pitch dX", 245760, 512, 136, "nn", cf)
shape("note bwd dh", 16384, 512, 1536, "nn", cf)
