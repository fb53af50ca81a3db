"""ctypes binding of ``libpolydis_b200.so`` (the C-ABI declared in ``include/polydis_b200.h``).

The library is the product's only compute path.  If it is missing (not built) this module raises
at import -- there is deliberately no PyTorch / CPU fallback.
"""
import ctypes
import os

from . import build as _build

_P, _L, _I, _F = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_float

# name -> argument ctypes (every function returns int: 0 ok, cudaError_t > 0, -22 bad argument)
SIGNATURES = {
    "pd_gemm_f32": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _P],
    "pd_gemm_tf32": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _P],
    "pd_gemm_tf32_cfg": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _I, _P],
    "pd_gemm_bf16": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _P],
    "pd_f32_to_bf16": [_P, _L, _L, _I, _P, _L, _P],
    "pd_tf32_split": [_P, _L, _L, _I, _P, _P, _L, _P],
    "pd_tf32_split3": [_P, _L, _L, _I, _P, _L, _I, _P],
    "pd_colsum_f32": [_P, _L, _I, _I, _P, _I, _P],
    "pd_colsum_seq_f32": [_P, _L, _I, _I, _I, _P, _P, _I, _P],
    "pd_sum_steps_f32": [_P, _L, _L, _I, _P, _L, _L, _I, _P],
    "pd_gru_gates_fwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _P],
    "pd_gru_gates_fwd_split3": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _P, _L, _P],
    "pd_gru_gates_bwd": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I,
                         _I, _I, _P],
    "pd_gru_step_tf32": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _P],
    "pd_gru_step_tma": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "pd_gru_step_tma3": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "pd_gru_step_tmax": [_P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "pd_gru_step_tma3x": [_P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "pd_gru_step_tma_variant": [_I],
    "pd_gru128_fwd": [_P, _L, _L, _P, _P, _P, _P, _L, _L, _P, _L, _L, _P, _L, _L, _L, _I, _I, _I, _P],
    "pd_gru128_fwd_perm": [_P, _L, _L, _P, _P, _P, _P, _L, _L, _P, _L, _L, _P, _L, _L, _L, _I, _I, _I, _P, _P],
    "pd_gru128_bwd": [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _P, _P, _L, _L, _P, _L, _L, _L, _I, _I, _P],
    "pd_greedy_decode_small": [_I] + [_P] * 3 + [_L] + [_P] * 6 + [_L] + [_P] * 27 + [_P],
    "pd_grid_prepare": [_P, _L, _P, _P, _P, _P, _P],
    "pd_prmat_to_grid": [_P, _L, _P, _P, _P],
    "pd_grid_to_prmat": [_P, _L, _P, _P],
    "pd_pack_tokens": [_P, _L, _P, _P],
    "pd_roll_prmat": [_P, _P, _L, _P, _P],
    "pd_expand_chord": [_P, _P, _L, _I, _P, _P],
    "pd_slerp_path": [_P, _P, _I, _I, _I, _P, _P],
    "pd_note_embed_fwd": [_P, _L, _P, _P, _P, _L, _P],
    "pd_note_embed_bwd": [_P, _L, _P, _L, _P, _P, _P],
    "pd_greedy_pick": [_P, _L, _P, _L, _L, _I, _P, _L, _P, _P],
    "pd_greedy_pick_embed": [_P, _L, _P, _L, _L, _I, _P, _L, _P, _P, _P, _P, _L, _P],
    "pd_dur_token": [_P, _L, _L, _P, _P],
    "pd_dur_decode_fwd": [_P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "pd_dur_decode_bwd": [_P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    "pd_transpose_f32": [_P, _I, _I, _P, _P],
    "pd_chord_feedback": [_P, _L, _P, _L, _P, _L, _I, _P, _P, _L, _P],
    "pd_chord_targets": [_P, _I, _P, _P, _P, _P],
    "pd_texture_frontend_fwd": [_P, _P, _P, _I, _I, _P, _P],
    "pd_texture_frontend_bwd": [_P, _P, _P, _I, _I, _P, _P, _P, _P],
    "pd_ce_fwd": [_P, _L, _P, _L, _I, _I, _P, _P, _P],
    "pd_ce_bwd": [_P, _L, _P, _L, _I, _I, _P, _P, _P, _L, _P],
    "pd_exp_fwd": [_P, _L, _P, _P],
    "pd_mul_f32": [_P, _P, _L, _P, _P],
    "pd_add_f32": [_P, _P, _L, _P, _P],
    "pd_select_rows": [_P, _L, _P, _L, _P, _P, _L, _L, _I, _P],
    "pd_select_rows_bwd": [_P, _L, _P, _P, _L, _P, _L, _L, _I, _P],
    "pd_reparam_fwd": [_P, _P, _P, _I, _I, _P, _L, _P],
    "pd_reparam_bwd": [_P, _L, _P, _I, _I, _P, _P, _P],
    "pd_kl_fwd": [_P, _P, _L, _P, _P],
    "pd_kl_bwd": [_P, _P, _L, _P, _P, _P, _P],
    "pd_sumsq_f32": [_P, _L, _P, _P],
    "pd_counter_inc": [_P, _P],
    "pd_adam_clip_step": [_P, _P, _P, _P, _L, _P, _P, _F, _F, _F, _F, _F, _F, _F, _P],
    # packed note level (length-sorted rows, slot-major buffers, device live-row table)
    "pd_pack_order": [_P, _I, _P, _P, _P, _P],
    "pd_pack_grid": [_P, _P, _P, _I, _P, _P, _P, _P, _P],
    "pd_gather_rows_f32": [_P, _L, _P, _L, _I, _P, _L, _P],
    "pd_sum_slots_rows_f32": [_P, _L, _L, _I, _P, _P, _L, _L, _I, _P],
    "pd_colsum_rows_f32": [_P, _L, _L, _I, _P, _I, _P, _I, _P],
    "pd_gemm_tf32_rows": [_P, _L, _L, _P, _L, _L, _P, _L, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "pd_gru_step_tmax_rows": [_P, _L, _P, _L, _P, _L, _P, _L, _I, _P, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P, _P],
    "pd_gru_gates_bwd_rows": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P, _P],
    "pd_dur_decode_fwd_rows": [_P, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "pd_dur_decode_bwd_rows": [_P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _I, _P],
    "pd_note_embed_bwd_rows": [_P, _L, _P, _L, _P, _P, _P, _I, _P],
    "pd_gru128_bwd_rows": [_P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _L, _L, _P, _P, _P, _L, _L, _P, _L, _L, _L, _I, _I, _P, _I, _P],
    "pd_gru_gates_bwd_z": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _P, _L, _P],
    "pd_texture_frontend_fwd_ix": [_P, _P, _P, _I, _I, _P, _P, _P],
    "pd_texture_frontend_bwd_ix": [_P, _P, _I, _I, _P, _P, _P, _P],
    "pd_gru_step_tma_bf16": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P],
    "pd_gru_step_tma_bf16_units": [_P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _P],
    "pd_gru_gates_bwd_zb": [_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _I, _P, _L, _P,
                            _L, _P],
    "pd_set_pdl": [_I],
    "pd_dur_quad_max_notes": [_I],
    "pd_ar_flag_bytes": [_I],                      # (returns bytes, not a status)
    "pd_ar_limit": [_I],                           # (returns the limit, not a status)
    "pd_ipc_alloc": [_L, _P, _P],
    "pd_ipc_open": [_P, _P],
    "pd_ipc_close": [_P],
    "pd_ipc_free": [_P],
    "pd_allreduce_p2p": [_P, _I, _I, _L, _L, _L, _F, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P],
    "pd_ar_norm_total": [_P, _I, _I, _I, _P, _P],
    "pd_gemm_tf32_splits": [_I, _I, _I],          # (returns the split count, not a status: call through ``lib``)
}

LIB_PATH = _build.LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built; run `python -m polydis_b200.build` (needs nvcc). "
            "polydis_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the symbol is missing
        fn.argtypes = args
        fn.restype = ctypes.c_int
    return lib


lib = _load()
GREEDY_SMALL_WS_FLOATS = 310368      # PD_GREEDY_SMALL_WS_FLOATS of include/polydis_b200.h

#: incremented on every kernel-launching library call (bench.py reports it as ``gpu_launches``)
call_count = 0


def check(name, code):
    if code != 0:
        raise RuntimeError(f"libpolydis_b200: {name} failed with code {code}"
                           + (" (bad argument)" if code == -22 else " (cudaError_t)"))


def call(name, *args):
    global call_count
    call_count += 1
    code = getattr(lib, name)(*args)
    if code != 0:
        check(name, code)
