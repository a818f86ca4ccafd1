"""ctypes binding of libes_b200.so (include/es_b200.h).  The structures mirror the header field
for field; tests/test_cabi_cpu.py parses the header and checks that they still agree."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import build as _build

ES_ABI_VERSION = 14
ES_GATHER_MATERIALIZE, ES_GATHER_FUSED = 1, 2
ES_MAX_ENC_BLOCKS = 2
ES_MAX_DEC_LAYERS = 24
ES_MAX_DEC_BLOCKS = 8
KERNEL_KINDS = ["embed", "enc_gemm", "attention", "fuse", "predictor", "variance", "lenreg",
                "dec_proj", "dec_layer", "mel", "poolmask", "phoneme"]

_fp = C.c_void_p      # device pointers travel as integers


class es_config_t(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "embed_dim", "dim", "kernel_size", "head", "expansion", "n_blocks", "block_depth",
        "decoder_kernel_size", "n_mel", "n_symbols")]


class es_enc_block_w_t(C.Structure):
    _fields_ = [(n, _fp) for n in (
        "merge_w", "qkv_w", "proj_w", "proj_b", "ln1_g", "ln1_b", "ffn1_w", "ffn1_tapb", "ffn1_b",
        "ffn2_w", "ffn2_b", "ln2_g", "ln2_b", "merge_w_h16", "qkv_w_h16", "proj_w_h16", "ffn1_w_h16", "ffn2_w_h16",
        "merge2_w", "merge2_w_h16")]


class es_predictor_w_t(C.Structure):
    _fields_ = [(n, _fp) for n in (
        "conv1_w", "conv1_b", "ln1_g", "ln1_b", "conv2_w", "conv2_b", "ln2_g", "ln2_b", "lin_w", "lin_b",
        "bins", "table", "conv1_w_h16", "conv2_w_h16")]


class es_dec_layer_w_t(C.Structure):
    _fields_ = [(n, _fp) for n in ("dw_w", "dw_b", "pw_w", "pw_b", "ln_g", "ln_b", "pw_w_h16")]


class es_weights_t(C.Structure):
    _fields_ = [
        ("enc", es_enc_block_w_t * ES_MAX_ENC_BLOCKS),
        ("fuse_a0", _fp), ("fuse_g", _fp), ("fuse_gb", _fp), ("fuse_c", _fp),
        ("fuse_u_h16", _fp), ("fuse_a0_h16", _fp),
        ("pitch", es_predictor_w_t), ("energy", es_predictor_w_t), ("duration", es_predictor_w_t),
        ("dproj_w", _fp), ("dproj_b", _fp), ("dproj_ln_g", _fp), ("dproj_ln_b", _fp),
        ("dec", es_dec_layer_w_t * ES_MAX_DEC_LAYERS),
        ("blk_ln_g", _fp * ES_MAX_DEC_BLOCKS), ("blk_ln_b", _fp * ES_MAX_DEC_BLOCKS),
        ("mel_w", _fp), ("mel_b", _fp), ("dproj_w_h16", _fp), ("mel_w_h16", _fp),
    ]


ES_HG_MAX_UPS = 6
ES_HG_MAX_RES = 4


class es_hifigan_config_t(C.Structure):
    _fields_ = [("n_mel", C.c_int32), ("initial_channel", C.c_int32), ("n_up", C.c_int32),
                ("up_rate", C.c_int32 * ES_HG_MAX_UPS), ("up_kernel", C.c_int32 * ES_HG_MAX_UPS), ("n_res", C.c_int32),
                ("res_kernel", C.c_int32 * ES_HG_MAX_RES), ("res_dilation", (C.c_int32 * 3) * ES_HG_MAX_RES)]


class es_hg_conv_w_t(C.Structure):
    _fields_ = [("w", _fp), ("b", _fp)]


class es_hg_resblock_w_t(C.Structure):
    _fields_ = [("convs1", es_hg_conv_w_t * 3), ("convs2", es_hg_conv_w_t * 3)]


class es_hifigan_weights_t(C.Structure):
    _fields_ = [("conv_pre", es_hg_conv_w_t), ("ups", es_hg_conv_w_t * ES_HG_MAX_UPS),
                ("res", (es_hg_resblock_w_t * ES_HG_MAX_RES) * ES_HG_MAX_UPS), ("conv_post", es_hg_conv_w_t)]


# name -> (restype, argtypes); exactly the functions include/es_b200.h declares
_i, _sz, _vp, _ll = C.c_int, C.c_size_t, C.c_void_p, C.c_longlong
PROTOTYPES = {
    "es_abi_version": (_i, []),
    "es_last_error": (C.c_char_p, []),
    "es_model_create": (_i, [C.POINTER(es_config_t), C.POINTER(es_weights_t), C.POINTER(_vp)]),
    "es_model_destroy": (None, [_vp]),
    "es_model_set_tensor_core": (_i, [_vp, _i]),
    "es_model_set_decoder_gather": (_i, [_vp, _i]),
    "es_model_set_fused_phoneme": (_i, [_vp, _i]),
    "es_model_set_ragged_schedule": (_i, [_vp, _i]),
    "es_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "es_encoder_forward": (_i, [_vp, _vp, _i, _i] + [_vp] * 12 + [_vp, _sz]),
    "es_length_regulate": (_i, [_vp, _vp, _i, _i, _i] + [_vp] * 6),
    "es_frame_rows": (_i, [_vp, _vp, _i, _i, _i] + [_vp] * 3),
    "es_collate": (_i, [_vp, _i, _i] + [_vp] * 13),
    "es_mel_to_half": (_i, [_vp, _vp, _vp, _sz]),
    "es_decoder_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _sz]),
    "es_decoder_forward_gathered": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz]),
    "es_dense_layout": (_i, [_i, _i, _i, _i]),
    "es_check_async_errors": (_i, [_vp]),
    "es_debug_set_trace": (_i, [_vp]),
    "es_debug_set_phoneme_trace": (_i, [_vp]),
    "es_launch_count": (C.c_uint64, []),
    "es_profile_begin": (_i, [_i]),
    "es_profile_end": (_i, []),
    "es_profile_collect": (_i, [_vp, _vp, _i, C.POINTER(_i)]),
    "es_hifigan_create": (_i, [C.POINTER(es_hifigan_config_t), C.POINTER(es_hifigan_weights_t), C.POINTER(_vp)]),
    "es_hifigan_destroy": (None, [_vp]),
    "es_hifigan_workspace_bytes": (_sz, [_vp, _i, _i]),
    "es_hifigan_forward": (_i, [_vp, _vp, _i, _i, _vp, C.c_longlong, C.c_longlong, C.c_longlong, _vp, _vp, _sz]),
    "es_loss_workspace_bytes": (_sz, []),
    "es_loss": (_i, [_vp, _i, _i, _i, _i] + [_vp] * 15 + [_vp, _sz]),
    "es_adamw_step": (_i, [_vp, _sz, _vp, _vp, _vp, _vp] + [C.c_float] * 7),
    "es_t_gemm": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _ll, _i, _vp, _i, _ll, _i, _vp, _i, _ll, _vp, _i, _i, _i]),
    "es_t_set_tensor_core": (None, [_i]),
    "es_t_im2col": (_i, [_vp, _vp, _vp] + [_i] * 7),
    "es_t_col2im": (_i, [_vp, _vp, _vp] + [_i] * 8),
    "es_t_dwconv_fwd": (_i, [_vp] * 5 + [_i] * 4),
    "es_t_dwconv_bwd_workspace_floats": (_sz, [_i] * 4),
    "es_t_dwconv_bwd": (_i, [_vp] * 7 + [_i] * 4 + [_vp, _sz]),
    "es_t_layernorm_fwd": (_i, [_vp] * 7 + [_ll, _i]),
    "es_t_layernorm_bwd_workspace_floats": (_sz, [_ll, _i]),
    "es_t_layernorm_bwd": (_i, [_vp] * 8 + [_ll, _i, _vp, _sz]),
    "es_t_colsum_workspace_floats": (_sz, [_ll, _i]),
    "es_t_colsum": (_i, [_vp] * 4 + [_ll, _i, _i, _vp, _sz]),
    "es_t_act_fwd": (_i, [_vp, _vp, _vp, _ll, _i]),
    "es_t_act_bwd": (_i, [_vp, _vp, _vp, _vp, _ll, _i]),
    "es_t_softmax_fwd": (_i, [_vp, _vp, _vp, _ll, _i, C.c_float]),
    "es_t_softmax_bwd": (_i, [_vp, _vp, _vp, _vp, _ll, _i, C.c_float]),
    "es_t_gather_rows": (_i, [_vp] * 4 + [_ll, _i]),
    "es_t_scatter_add_rows": (_i, [_vp] * 4 + [_ll, _i, _i]),
    "es_t_expand_rows": (_i, [_vp] * 4 + [_i] * 4),
    "es_t_reduce_rows": (_i, [_vp] * 4 + [_i] * 4),
    "es_t_bucketize": (_i, [_vp, _vp, _vp, _i, _vp, _ll]),
    "es_t_axpby": (_i, [_vp] * 4 + [_ll, C.c_float, C.c_float]),
    "es_t_mask_rows": (_i, [_vp] * 4 + [_ll, _i]),
    "es_t_copy2d": (_i, [_vp, _vp, _i, _vp, _i, _ll, _i, _i]),
    "es_selftest_umma_gemm": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "es_selftest_attention": (_i, [_vp, _i, _i, _i, _i, C.c_float, _vp, _vp, _i]),
}

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    # $ES_B200_LIB: an alternative build of the same sources (A/B experiments, efficientspeech_b200/build.py)
    return os.environ.get("ES_B200_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library.  There is no fallback: a missing/unloadable library raises."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.isfile(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: run `python -m efficientspeech_b200.build` "
                               "(there is no CPU/PyTorch fallback for the acoustic path)")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    got = lib.es_abi_version()
    if got != ES_ABI_VERSION:
        raise RuntimeError(f"libes_b200.so ABI {got} != binding ABI {ES_ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().es_last_error()
        raise RuntimeError("es_b200: " + (msg.decode() if msg else f"error {rc}"))
