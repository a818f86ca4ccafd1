"""CPU-side checks of the boundary: header <-> ctypes agreement, exported symbols, state-dict
layout, failure behaviour without a GPU.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import efficientspeech_b200 as es
from efficientspeech_b200 import _cabi
from efficientspeech_b200.config import VARIANTS
from efficientspeech_b200.params import init_state_dict, param_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "es_b200.h")).read()


def _struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s_t;" % (name, name), HEADER, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        out.append(re.sub(r"\[.*\]", "", decl.split()[-1].lstrip("*")))
    return out


def test_ctypes_structs_match_header():
    for name, cls in [("es_config", _cabi.es_config_t), ("es_enc_block_w", _cabi.es_enc_block_w_t),
                      ("es_predictor_w", _cabi.es_predictor_w_t), ("es_dec_layer_w", _cabi.es_dec_layer_w_t),
                      ("es_weights", _cabi.es_weights_t), ("es_hifigan_config", _cabi.es_hifigan_config_t),
                      ("es_hg_conv_w", _cabi.es_hg_conv_w_t), ("es_hg_resblock_w", _cabi.es_hg_resblock_w_t),
                      ("es_hifigan_weights", _cabi.es_hifigan_weights_t)]:
        assert _struct_fields(name) == [f[0] for f in cls._fields_], name
    for macro, val in [("ES_ABI_VERSION", _cabi.ES_ABI_VERSION), ("ES_MAX_DEC_LAYERS", _cabi.ES_MAX_DEC_LAYERS),
                       ("ES_MAX_DEC_BLOCKS", _cabi.ES_MAX_DEC_BLOCKS), ("ES_MAX_ENC_BLOCKS", _cabi.ES_MAX_ENC_BLOCKS),
                       ("ES_HG_MAX_UPS", _cabi.ES_HG_MAX_UPS), ("ES_HG_MAX_RES", _cabi.ES_HG_MAX_RES)]:
        assert int(re.search(r"#define %s (\d+)" % macro, HEADER).group(1)) == val


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = _cabi.load()
    declared = set(re.findall(r"\b(es_[a-z_0-9]+)\s*\(", re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)))
    assert declared == set(_cabi.PROTOTYPES), declared ^ set(_cabi.PROTOTYPES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.es_abi_version() == _cabi.ES_ABI_VERSION
    assert isinstance(lib.es_launch_count(), int)


@pytest.mark.parametrize("vname", list(VARIANTS))
def test_state_dict_layout_equals_reference_layout(vname):
    cfg = VARIANTS[vname]
    model = es.build_model(vname)
    sd = model.state_dict()
    want = param_shapes(cfg)
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == want[k], k
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == {"tiny": 266417, "small": 952465, "base": 3953489}[vname]
    es.load_numpy_state(model, init_state_dict(cfg, seed=3))                      # strict=True round trip


def test_state_dict_and_default_init_match_reference_when_present():
    from oracle.ref_shim import reference_available, import_reference_layers
    if not reference_available():
        pytest.skip("reference tree not mounted")
    layers = import_reference_layers()
    cfg = VARIANTS["small"]
    torch.manual_seed(7)
    ref = layers.Phoneme2Mel(
        layers.PhonemeEncoder(pitch_stats=cfg.pitch_stats, energy_stats=cfg.energy_stats, reduction=2),
        layers.MelDecoder(dim=64, kernel_size=5, n_blocks=3, block_depth=2))
    torch.manual_seed(7)
    ours = es.build_model("small")
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k          # same names, shapes AND default initialisation
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict(ref.state_dict(), strict=True)


def test_no_cpu_fallback():
    model = es.build_model("tiny")
    x = {"phoneme": torch.ones(2, 8, dtype=torch.int32), "phoneme_mask": torch.zeros(2, 8, dtype=torch.bool)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.decoder(torch.zeros(1, 4, 128))


def test_compute_entry_fails_cleanly_without_device():
    if torch.cuda.is_available():
        pytest.skip("device present")
    lib = _cabi.load()
    cfg = _cabi.es_config_t(128, 32, 3, 1, 1, 2, 2, 5, 80, 153)
    w = _cabi.es_weights_t()
    h = ctypes.c_void_p(None)
    assert lib.es_model_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(h)) == 0
    assert lib.es_workspace_bytes(h, 4, 16, 64) > 0
    # invalid arguments are rejected before any launch
    assert lib.es_decoder_forward(h, None, 0, 4, None, None, None, 0) != 0
    assert b"es_decoder_forward" in lib.es_last_error()
    lib.es_model_destroy(h)
    bad = _cabi.es_config_t(128, 48, 3, 1, 1, 2, 2, 5, 80, 153)
    assert lib.es_model_create(ctypes.byref(bad), ctypes.byref(w), ctypes.byref(h)) != 0


def test_path_switches_and_workspace_planning_on_the_host():
    """Host logic of the C ABI that needs no device: the A/B switches validate their argument, and the workspace
    plan is monotone and covers the gathered decoder entry (projection table + frame -> row map)."""
    lib = _cabi.load()
    cfg = _cabi.es_config_t(128, 32, 3, 1, 1, 2, 2, 5, 80, 153)
    w = _cabi.es_weights_t()
    h = ctypes.c_void_p(None)
    assert lib.es_model_create(ctypes.byref(cfg), ctypes.byref(w), ctypes.byref(h)) == 0
    try:
        for mode in (_cabi.ES_GATHER_MATERIALIZE, _cabi.ES_GATHER_FUSED):
            assert lib.es_model_set_decoder_gather(h, mode) == 0
        assert lib.es_model_set_decoder_gather(h, 7) != 0 and lib.es_model_set_decoder_gather(h, 0) != 0
        assert b"gather mode" in lib.es_last_error()
        assert lib.es_model_set_fused_phoneme(h, 0) == 0 and lib.es_model_set_fused_phoneme(h, 1) == 0
        assert lib.es_model_set_tensor_core(h, 0) == 0 and lib.es_model_set_tensor_core(h, 1) == 0
        B, N, T, dx2 = 8, 128, 768, 128
        enc = lib.es_workspace_bytes(h, B, N, 0)
        dec = lib.es_workspace_bytes(h, B, 0, T)
        both = lib.es_workspace_bytes(h, B, N, T)
        assert enc > 0 and dec >= 3 * B * T * dx2 * 4                  # three rotating [B,T,dx2] fp32 buffers
        assert both >= enc + dec                                       # one plan after the other
        assert both - enc - dec >= (B * N + 1) * dx2 * 4 + B * T * 4 - 1024   # + projection table + int32 row map
        assert lib.es_workspace_bytes(h, 2 * B, N, T) > both
        assert lib.es_workspace_bytes(None, B, N, T) == 0
        # argument checks come before any device work
        assert lib.es_frame_rows(h, None, B, N, T, None, None, None) != 0
        assert b"es_frame_rows" in lib.es_last_error()
        assert lib.es_decoder_forward_gathered(h, None, B, N, T, None, None, None, 1, None, None, 0) != 0
    finally:
        lib.es_model_destroy(h)
