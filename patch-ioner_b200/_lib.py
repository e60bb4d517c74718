"""ctypes binding of libpio_sm100.so (the C ABI declared in include/pio.h).

There is no CPU fallback: if the library is missing this module raises, and every compute entry
point fails with PIO_ECUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpio_sm100.so")

PIO_FP32, PIO_BF16 = 0, 1
DT_F32, DT_BF16, DT_I32 = 0, 1, 2
POOL_MEAN, POOL_GAUSS, POOL_ATTN = 0, 1, 2
ACT_NONE, ACT_GELU_ERF, ACT_GELU_NEW, ACT_TANH, ACT_RELU = 0, 1, 2, 3, 4

_fp = C.c_void_p  # device pointers travel as integers


class PioLinear(C.Structure):
    _fields_ = [("A", _fp), ("W", _fp), ("C", _fp), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("lda", C.c_int), ("ldw", C.c_int), ("ldc", C.c_int), ("a_dt", C.c_int), ("c_dt", C.c_int),
                ("bias", _fp), ("colscale", _fp), ("gamma", _fp), ("residual", _fp), ("res_rowscale", _fp),
                ("ldres", C.c_int), ("alpha", C.c_float), ("act", C.c_int),
                ("rows_per_group", C.c_int), ("group_stride", C.c_int), ("group_offset", C.c_int),
                ("argmax_val", _fp), ("argmax_idx", _fp), ("argmax_sumexp", _fp), ("argmax_ld", C.c_int),
                ("exp_ref", _fp), ("exp_psum", _fp), ("exp_pmax", _fp), ("exp_ld", C.c_int), ("w_static", C.c_int)]


VIT_BLOCK_FIELDS = ["ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
                    "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2"]


class PioVitBlock(C.Structure):
    _fields_ = [(n, _fp) for n in VIT_BLOCK_FIELDS]


class PioVitWeights(C.Structure):
    _fields_ = [("cls_token", _fp), ("register_tokens", _fp), ("patch_w", _fp), ("patch_b", _fp),
                ("blk", PioVitBlock * 12), ("norm_w", _fp), ("norm_b", _fp)]


GPT_BLOCK_FIELDS = ["ln1_w", "ln1_b", "attn_w", "attn_b", "proj_w", "proj_b",
                    "ln2_w", "ln2_b", "fc_w", "fc_b", "fc2_w", "fc2_b"]


class PioGptBlock(C.Structure):
    _fields_ = [(n, _fp) for n in GPT_BLOCK_FIELDS]


class PioDecoderWeights(C.Structure):
    _fields_ = [("wte", _fp), ("wpe", _fp), ("blk", PioGptBlock * 4), ("lnf_w", _fp), ("lnf_b", _fp),
                ("prefix_w", _fp), ("prefix_b", _fp), ("prefix_size", C.c_int)]


class PioGpt2Weights(C.Structure):
    _fields_ = [("wte", _fp), ("wpe", _fp), ("lnf_w", _fp), ("lnf_b", _fp), ("blk", C.POINTER(PioGptBlock)),
                ("n_layer", C.c_int), ("n_head", C.c_int)]


MAPPER_LAYER_FIELDS = ["norm1_w", "norm1_b", "q_w", "kv_w", "proj_w", "proj_b",
                       "norm2_w", "norm2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b"]


class PioMapperLayer(C.Structure):
    _fields_ = [(n, _fp) for n in MAPPER_LAYER_FIELDS]


class PioMapperWeights(C.Structure):
    _fields_ = [("clip_size", C.c_int), ("project_len", C.c_int), ("prefix_len", C.c_int), ("n_layer", C.c_int),
                ("n_head", C.c_int), ("hidden", C.c_int), ("linear_w", _fp), ("linear_b", _fp), ("prefix_const", _fp),
                ("layers", C.POINTER(PioMapperLayer))]


# name -> (restype, argtypes); must list every symbol include/pio.h declares (tests check this)
SIGNATURES = {
    "pio_last_error": (C.c_char_p, []),
    "pio_version": (C.c_int, []),
    "pio_launch_count": (C.c_longlong, []),
    "pio_reset_launch_count": (None, []),
    "pio_release_scratch": (None, []),
    "pio_linear": (C.c_int, [C.POINTER(PioLinear), C.c_int, _fp]),
    "pio_argmax_slabs": (C.c_int, [C.c_int, C.c_int]),
    "pio_argmax_finish": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _fp, _fp]),
    "pio_layernorm": (C.c_int, [_fp, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "pio_vit_create": (C.c_int, [C.POINTER(_fp), C.POINTER(PioVitWeights), C.c_int, _fp]),
    "pio_vit_destroy": (None, [_fp]),
    "pio_vit_workspace_bytes": (C.c_size_t, [_fp, C.c_int, C.c_int]),
    "pio_vit_forward": (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_attention_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pio_vit_attention": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_size_t, _fp]),
    "pio_cls_attention": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp]),
    "pio_vit_block_workspace_bytes": (C.c_size_t, [_fp, C.c_int]),
    "pio_vit_block_rows": (C.c_int, [_fp, C.c_int, _fp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, _fp, C.c_size_t, _fp]),
    "pio_gather_rows": (C.c_int, [_fp, C.c_longlong, _fp, C.c_int, C.c_int, _fp, _fp]),
    "pio_segment_mean": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp]),
    "pio_preprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "pio_preprocess": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), _fp, _fp, C.c_size_t, _fp]),
    "pio_ctx_clean": (C.c_int, [_fp, C.c_longlong, C.c_longlong, _fp, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_float, C.c_int, _fp, _fp]),
    "pio_cls_head_attention": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp]),
    "pio_pool_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "pio_pool_boxes": (C.c_int, [_fp, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_float, _fp, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_pool_grid": (C.c_int, [_fp, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_float, _fp, _fp]),
    "pio_trace_bins": (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp]),
    "pio_region_mean_weights": (C.c_int, [C.c_int, C.c_float, _fp, _fp]),
    "pio_bank_create": (C.c_int, [C.POINTER(_fp), _fp, C.c_longlong, C.c_int, C.c_int, _fp]),
    "pio_bank_destroy": (None, [_fp]),
    "pio_bank_rows": (C.c_longlong, [_fp]),
    "pio_project_workspace_bytes": (C.c_size_t, [_fp, C.c_int]),
    "pio_project": (C.c_int, [_fp, _fp, C.c_int, C.c_float, C.c_int, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_best_sims": (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_project_rescale": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp]),
    "pio_project_finish": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "pio_decoder_create": (C.c_int, [C.POINTER(_fp), C.POINTER(PioDecoderWeights), C.c_int, _fp]),
    "pio_decoder_destroy": (None, [_fp]),
    "pio_decode_workspace_bytes": (C.c_size_t, [_fp, C.c_int, C.c_int]),
    "pio_set_decode_fused": (C.c_int, [C.c_int, C.c_int]),
    "pio_decode_debug_layout": (C.c_int, [_fp, C.c_int, _fp, C.c_int]),
    "pio_decode_greedy": (C.c_int, [_fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_l2_normalize": (C.c_int, [_fp, C.c_int, C.c_int, _fp]),
    "pio_decoder_create_gpt2": (C.c_int, [C.POINTER(_fp), C.POINTER(PioGpt2Weights), C.c_int, _fp]),
    "pio_decode_prompt_workspace_bytes": (C.c_size_t, [_fp, C.c_int, C.c_int, C.c_int]),
    "pio_decode_greedy_prompt": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    "pio_decode_greedy_prompt_eos": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, C.POINTER(C.c_int), _fp,
                                               C.c_size_t, _fp]),
    "pio_decode_beam_workspace_bytes": (C.c_size_t, [_fp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pio_decode_beam_prompt": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp, _fp, _fp,
                                         C.POINTER(C.c_int), _fp, C.c_size_t, _fp]),
    "pio_gpt2_score_workspace_bytes": (C.c_size_t, [_fp, C.c_int, C.c_int]),
    "pio_gpt2_score_tokens": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "pio_mapper_create": (C.c_int, [C.POINTER(_fp), C.POINTER(PioMapperWeights), C.c_int, _fp]),
    "pio_mapper_destroy": (None, [_fp]),
    "pio_mapper_workspace_bytes": (C.c_size_t, [_fp, C.c_int]),
    "pio_mapper_forward": (C.c_int, [_fp, _fp, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "pio_detok_rows": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_longlong,
                                 _fp, _fp, _fp]),
    "pio_entity_topk": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _fp, _fp, _fp]),
}

_lib = None


class PioError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) the CUDA library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PioError(f"{LIB_PATH} is missing: run `python patch-ioner_b200/build.py` (needs nvcc). "
                           "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise PioError(f"libpio_sm100 error {rc}: {lib().pio_last_error().decode(errors='replace')}")
