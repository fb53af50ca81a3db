"""polydis_b200 -- Blackwell-native hot path of PolyDis (chord/texture-disentangling piano VAE).

Drop-in for the reference's ``model.py`` ``DisentangleVAE`` whose training forward/backward and
greedy decoding run on hand-written sm_100a CUDA kernels behind a C-ABI shared library
(``csrc/`` -> ``libpolydis_b200.so``, declared in ``include/polydis_b200.h``).  There is no CPU
fallback: importing the compute modules without the built library raises.
"""
__all__ = ["synth", "weights"]
