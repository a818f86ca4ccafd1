"""Host-side mirror of the reference's acoustic-model API, backed by libes_b200.so.

``PhonemeEncoder``, ``MelDecoder`` and ``Phoneme2Mel`` keep the constructor signatures,
attribute names, forward signatures and ``state_dict()`` layout of the reference
(layers/networks.py:261-434, exported at layers/__init__.py:1, built at model.py:132-147), so
``from efficientspeech_b200 import PhonemeEncoder, MelDecoder, Phoneme2Mel`` is a drop-in for
``from layers import ...``.  The torch sub-modules below only HOLD parameters (same names,
shapes and default initialisation as upstream); no torch op computes anything on the path.
Every forward packs the weights (once per weight version), then calls the C ABI on the
current CUDA stream.  There is no CPU fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import _cabi, packing
from .config import ESConfig, N_SYMBOLS

__all__ = ["PhonemeEncoder", "MelDecoder", "Phoneme2Mel", "GraphedForward", "Encoder", "Fuse", "AcousticDecoder",
           "FeatureUpsampler", "SelfAttention", "MixFFN", "mel_to_half"]


# ------------------------------------------------------------------------------------------
# parameter holders (state-dict compatible with the reference; never called as torch ops)
# ------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only holds parameters; call PhonemeEncoder / "
                           "MelDecoder / Phoneme2Mel, which run the fused sm_100a kernels")


class SelfAttention(_Holder):
    """Parameters of layers/blocks.py:32-41 (qkv without bias, proj with bias)."""

    def __init__(self, dim, num_heads=1, qkv_bias=False):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3 * num_heads, bias=qkv_bias)
        self.proj = nn.Linear(dim * num_heads, dim)


class MixFFN(_Holder):
    """Parameters of layers/blocks.py:8-19."""

    def __init__(self, dim, expansion_factor):
        super().__init__()
        hidden = dim * expansion_factor
        self.mlp1 = nn.Linear(dim, hidden)
        self.conv = nn.Conv1d(hidden, hidden, 3, padding=1)
        self.mlp2 = nn.Linear(hidden, dim)


class Encoder(_Holder):
    """Parameters of layers/networks.py:15-47."""

    def __init__(self, depth=2, embed_dim=128, kernel_size=3, expansion=1, reduction=4, head=1):
        super().__init__()
        cfg = ESConfig(depth=depth, embed_dim=embed_dim, kernel_size=kernel_size, expansion=expansion,
                       reduction=reduction, head=head)
        cfg.validate()
        self.dim_outs = cfg.enc_dims
        self.embed = nn.Embedding(N_SYMBOLS, embed_dim, padding_idx=0)
        self.attn_blocks = nn.ModuleList([])
        for cin, c, h, k, s in zip(cfg.enc_dims_in, cfg.enc_dims, cfg.enc_heads, cfg.enc_kernels, cfg.enc_strides):
            self.attn_blocks.append(nn.ModuleList([
                nn.Conv1d(cin, cin, kernel_size=k, stride=s, padding=k // 2, bias=False),
                nn.Conv1d(cin, c, kernel_size=1, bias=False),
                SelfAttention(c, num_heads=h),
                MixFFN(c, expansion),
                nn.LayerNorm(c),
                nn.LayerNorm(c)]))

    def get_feature_dims(self):
        return self.dim_outs


class Fuse(_Holder):
    """Parameters of layers/networks.py:168-187."""

    def __init__(self, dims, kernel_size=3):
        super().__init__()
        dim = dims[0]
        self.mlps = nn.ModuleList([])
        for d in dims:
            up = d // dim
            self.mlps.append(nn.ModuleList([
                nn.Linear(d, dim),
                nn.ConvTranspose1d(dim, dim, kernel_size=kernel_size, stride=up) if up > 1 else nn.Identity()]))
        self.fuse = nn.Linear(dim * len(dims), dim)


class FeatureUpsampler(_Holder):
    """The length regulator has no parameters (layers/networks.py:222-226)."""


class AcousticDecoder(_Holder):
    """Parameters of the pitch / energy / duration predictor (layers/networks.py:90-122)."""

    def __init__(self, dim, pitch_stats=None, energy_stats=None, n_mel_channels=80, duration=False):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        self.conv1 = nn.Sequential(nn.Conv1d(dim, dim, kernel_size=3, padding=1), nn.ReLU())
        self.norm1 = nn.LayerNorm(dim)
        self.conv2 = nn.Sequential(nn.Conv1d(dim, dim, kernel_size=3, padding=1), nn.ReLU())
        self.norm2 = nn.LayerNorm(dim)
        self.linear = nn.Linear(dim, 1)
        self.duration = duration
        if pitch_stats is not None:
            lo, hi = pitch_stats
            self.pitch_bins = nn.Parameter(torch.linspace(lo, hi, dim - 1), requires_grad=False)
            self.pitch_embedding = nn.Embedding(dim, dim)
        else:
            self.pitch_bins = None
            self.pitch_embedding = None
        if energy_stats is not None:
            lo, hi = energy_stats
            self.energy_bins = nn.Parameter(torch.linspace(lo, hi, dim - 1), requires_grad=False)
            self.energy_embedding = nn.Embedding(dim, dim)
        else:
            self.energy_bins = None
            self.energy_embedding = None


# ------------------------------------------------------------------------------------------
# device-side state shared by the public modules
# ------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"efficientspeech_b200: {what} must be a CUDA tensor -- the acoustic path "
                           "runs only as sm_100a kernels, there is no CPU fallback")


class _Backend:
    """Packed weights + es_model handle of one module ("encoder" or "decoder" half)."""

    def __init__(self, owner: nn.Module, cfg: ESConfig, part: str):
        self.owner, self.cfg, self.part = owner, cfg, part
        self.key = None
        self.flat: Optional[torch.Tensor] = None
        self.handle = C.c_void_p(None)
        self.workspace: Optional[torch.Tensor] = None
        self.tensor_core = True
        # ES_DEC_GATHER_MODE (1/2) overrides the default for A/B runs of the same command line
        self.gather_mode = int(os.environ.get("ES_DEC_GATHER_MODE", _cabi.ES_GATHER_FUSED))
        self.fused_phoneme = os.environ.get("ES_FUSED_PHONEME", "1") != "0"
        self.ragged_schedule = os.environ.get("ES_RAGGED_SCHEDULE", "1") != "0"

    def __del__(self):
        try:
            if self.handle.value:
                _cabi.load().es_model_destroy(self.handle)
        except Exception:
            pass

    def _state(self) -> Dict[str, np.ndarray]:
        prefix = self.part + "."
        return {prefix + k: v.detach().to("cpu", torch.float32).numpy()
                for k, v in self.owner.state_dict().items()}

    def ensure(self, device: torch.device) -> C.c_void_p:
        params = list(self.owner.parameters())
        key = (str(device), self.tensor_core, self.gather_mode, (self.fused_phoneme, self.ragged_schedule)) + \
            tuple((p.data_ptr(), p._version) for p in params)
        if key == self.key:
            return self.handle
        lib = _cabi.load()
        sd = self._state()
        folded = packing.fold_encoder(sd, self.cfg) if self.part == "encoder" else packing.fold_decoder(sd, self.cfg)
        if self.part == "encoder":
            # B operands of the tcgen05 row GEMMs; the library says which image format each layer uses
            # (resident taps image for es_umma_enc.cu, streamed units for es_umma_wide.cu)
            c = self.cfg

            def image(name: str, n: int, stride: int = 1) -> np.ndarray:
                w = folded[name]                                  # [taps][K][N_padded]
                lay = lib.es_dense_layout(int(w.shape[1]), int(n), int(w.shape[0]), int(stride))
                if lay >= 2:
                    return packing.canon_split_units(w, n, 128 if lay == 2 else 256)
                return packing.canon_split_taps(w, n)

            for i in range(2):
                ci, hi, hci = c.enc_dims[i], c.enc_heads[i], c.enc_dims[i] * c.expansion
                if i == 1:
                    folded["enc1.merge_w_h16"] = image("enc1.merge_w", ci, stride=2)
                    wm = folded["enc1.merge_w"]                     # [k'][Cin][C_padded]
                    if wm.shape[0] == 3 and lib.es_dense_layout(int(wm.shape[1]), int(ci), 3, 2) == 0 \
                            and lib.es_dense_layout(2 * int(wm.shape[1]), int(ci), 3, 1) >= 2:
                        # no strided tensor-core form for 3 taps (base): the same conv over paired rows (es_b200.h)
                        folded["enc1.merge2_w"] = packing.pair_stride2_taps(wm)
                        folded["enc1.merge2_w_h16"] = image("enc1.merge2_w", ci)
                folded[f"enc{i}.qkv_w_h16"] = image(f"enc{i}.qkv_w", 3 * hi * ci)
                folded[f"enc{i}.proj_w_h16"] = image(f"enc{i}.proj_w", ci)
                folded[f"enc{i}.ffn1_w_h16"] = image(f"enc{i}.ffn1_w", hci)
                folded[f"enc{i}.ffn2_w_h16"] = image(f"enc{i}.ffn2_w", ci)
            for which in ("pitch", "energy", "duration"):
                for cv in ("conv1", "conv2"):
                    folded[f"{which}.{cv}_w_h16"] = image(f"{which}.{cv}_w", c.dim)
            # Fuse as two tensor-core GEMMs (es_api.cu: fuse): all k transposed-conv taps at once
            k = folded["fuse_g"].shape[0]
            gcat = np.concatenate([folded["fuse_g"][t] for t in range(k)], axis=1)[None]       # [1][2d][k*d]
            if c.dim % 128 == 0:                                        # streamed units (es_umma_wide.cu)
                folded["fuse_u_h16"] = packing.canon_split_units(gcat, k * c.dim, 128)
                folded["fuse_a0_h16"] = packing.canon_split_units(folded["fuse_a0"][None], c.dim, 128)
            elif lib.es_dense_layout(2 * c.dim, k * c.dim, 1, 1) == 1 and lib.es_dense_layout(c.dim, c.dim, 1, 1) == 1:
                folded["fuse_u_h16"] = packing.canon_split_taps(gcat, k * c.dim)       # resident image (es_umma_enc.cu)
                folded["fuse_a0_h16"] = packing.canon_split_taps(folded["fuse_a0"][None], c.dim)
        if self.part == "decoder":
            # B operands of the tcgen05 kernels: W as [N][K], split into fp16 hi/lo, canonical order
            # dx2 == 128: whole-matrix image (weights stay resident in shared memory, es_umma_dec.cu);
            # dx2 == 256: K-chunked image streamed per tile (es_umma_dec256.cu)
            split = packing.canon_split_fp16 if self.cfg.dx2 <= 128 else packing.canon_split_chunks
            for l in range(self.cfg.n_dec_layers):
                folded[f"dec{l}.pw_w_h16"] = split(folded[f"dec{l}.pw_w"][0].T)
            folded["dproj_w_h16"] = split(folded["dproj_w"][0].T)
            folded["mel_w_h16"] = split(folded["mel_w"][0].T[:self.cfg.n_mel])
        flat_np, off = packing.pack(folded)
        self.flat = torch.from_numpy(flat_np).to(device)
        base = self.flat.data_ptr()
        W = _cabi.es_weights_t()

        def at(name):
            return base + 4 * off[name]

        if self.part == "encoder":
            for i in range(2):
                for f, _ in _cabi.es_enc_block_w_t._fields_:
                    if f"enc{i}.{f}" in off:
                        setattr(W.enc[i], f, at(f"enc{i}.{f}"))
            for f in ("fuse_a0", "fuse_g", "fuse_gb", "fuse_c"):
                setattr(W, f, at(f))
            for f in ("fuse_u_h16", "fuse_a0_h16"):
                if f in off:
                    setattr(W, f, at(f))
            for which in ("pitch", "energy", "duration"):
                pw = getattr(W, which)
                for f, _ in _cabi.es_predictor_w_t._fields_:
                    if f in ("bins", "table") and which == "duration":
                        continue
                    setattr(pw, f, at(f"{which}.{f}"))
        else:
            for f in ("dproj_w", "dproj_b", "dproj_ln_g", "dproj_ln_b", "mel_w", "mel_b", "dproj_w_h16", "mel_w_h16"):
                setattr(W, f, at(f))
            for l in range(self.cfg.n_dec_layers):
                for f, _ in _cabi.es_dec_layer_w_t._fields_:
                    setattr(W.dec[l], f, at(f"dec{l}.{f}"))
            for b in range(self.cfg.n_blocks):
                W.blk_ln_g[b] = at(f"blk{b}.ln_g")
                W.blk_ln_b[b] = at(f"blk{b}.ln_b")
        c = self.cfg
        cc = _cabi.es_config_t(c.embed_dim, c.dim, c.kernel_size, c.head, c.expansion, c.n_blocks, c.block_depth,
                               c.decoder_kernel_size, c.n_mel, c.n_symbols)
        if self.handle.value:
            lib.es_model_destroy(self.handle)
            self.handle = C.c_void_p(None)
        h = C.c_void_p(None)
        _cabi.check(lib.es_model_create(C.byref(cc), C.byref(W), C.byref(h)))
        _cabi.check(lib.es_model_set_tensor_core(h, 1 if self.tensor_core else 0))
        _cabi.check(lib.es_model_set_decoder_gather(h, int(self.gather_mode)))
        _cabi.check(lib.es_model_set_fused_phoneme(h, 1 if self.fused_phoneme else 0))
        _cabi.check(lib.es_model_set_ragged_schedule(h, 1 if self.ragged_schedule else 0))
        self.handle = h
        self.key = key
        return h

    def scratch(self, device: torch.device, B: int, N: int, T: int):
        need = _cabi.load().es_workspace_bytes(self.handle, B, N, T)
        if self.workspace is None or self.workspace.device != device or self.workspace.numel() < need:
            self.workspace = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=device)
        return self.workspace.data_ptr(), self.workspace.numel()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def mel_to_half(mel: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Opt-in fp16 copy of a mel tensor (es_mel_to_half): halves the device -> host bytes of a consumer that pulls the
    mel over PCIe.  Not part of the reference path -- ``Phoneme2Mel`` keeps returning fp32."""
    _require_cuda(mel, "mel")
    mel = mel.contiguous()
    if mel.dtype != torch.float32:
        raise RuntimeError("mel_to_half expects the fp32 mel")
    if out is None:
        out = torch.empty(mel.shape, dtype=torch.float16, device=mel.device)
    with torch.cuda.device(mel.device):
        _cabi.check(_cabi.load().es_mel_to_half(_stream(mel.device), mel.data_ptr(), out.data_ptr(), mel.numel()))
    return out


# ------------------------------------------------------------------------------------------
# public modules
# ------------------------------------------------------------------------------------------
class MelDecoder(nn.Module):
    """Mel Spectrogram Decoder -- drop-in for layers/networks.py:261-304.

    ``forward(features [B,T,4*dim]) -> mel [B,T,n_mel_channels]``.  One fused kernel per layer:
    depthwise conv k5 -> 1x1 GEMM -> bias -> tanh -> LayerNorm (-> + skip -> LayerNorm at block
    ends), then the mel head.
    """

    def __init__(self, dim, kernel_size=5, n_mel_channels=80, n_blocks=2, block_depth=2):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        dim_x2 = min(4 * dim, 256)
        dim_x4 = 4 * dim
        padding = kernel_size // 2
        self.proj = nn.Sequential(nn.Linear(dim_x4, dim_x2), nn.Tanh(), nn.LayerNorm(dim_x2))
        self.blocks = nn.ModuleList([])
        for _ in range(n_blocks):
            conv = nn.ModuleList([])
            for _ in range(block_depth):
                conv.append(nn.ModuleList([nn.Sequential(
                    nn.Conv1d(dim_x2, dim_x2, groups=dim_x2, kernel_size=kernel_size, padding=padding),
                    nn.Conv1d(dim_x2, dim_x2, kernel_size=1),
                    nn.Tanh()), nn.LayerNorm(dim_x2)]))
            self.blocks.append(nn.ModuleList([conv, nn.LayerNorm(dim_x2)]))
        self.mel_linear = nn.Linear(dim_x2, n_mel_channels)
        # embed_dim/reduction only fix dim = embed_dim // reduction; any pair with that quotient works
        cfg = ESConfig(embed_dim=dim, reduction=1, n_blocks=n_blocks, block_depth=block_depth,
                       decoder_kernel_size=kernel_size, n_mel=n_mel_channels)
        cfg.validate()
        self._cfg = cfg
        self._backend = _Backend(self, cfg, "decoder")

    def set_tensor_core(self, enable: bool) -> None:
        """True (default): tcgen05 split-fp16 decoder layers; False: fp32 SIMT kernels."""
        self._backend.tensor_core = bool(enable)

    def set_ragged_schedule(self, enable: bool) -> None:
        """True (default): with padded frames zeroed (B > 1) the decoder only schedules tiles that can reach a valid frame
        (include/es_b200.h: es_model_set_ragged_schedule); False: every padded frame is computed, as the reference does."""
        self._backend.ragged_schedule = bool(enable)

    def set_gather_mode(self, mode: int) -> None:
        """How ``Phoneme2Mel`` joins the length regulator and this decoder (include/es_b200.h, ES_GATHER_*):
        1 projection per phoneme + row-gather kernel; 2 (default) projection per phoneme, the first block gathers the
        rows itself.  Bit-identical results."""
        if mode not in (1, 2):
            raise ValueError("gather mode must be 1 or 2")
        self._backend.gather_mode = int(mode)

    def forward(self, features):
        _require_cuda(features, "features")
        B, T, Cin = features.shape
        if Cin != self._cfg.dx4:
            raise RuntimeError(f"features must have {self._cfg.dx4} channels, got {Cin}")
        dev = features.device
        with torch.cuda.device(dev):
            h = self._backend.ensure(dev)
            feats = features.contiguous().float()
            mel = torch.empty(B, T, self.n_mel_channels, dtype=torch.float32, device=dev)
            ws, wsn = self._backend.scratch(dev, B, 0, T)
            _cabi.check(_cabi.load().es_decoder_forward(h, _stream(dev), B, T, feats.data_ptr(), mel.data_ptr(), ws, wsn))
        return mel

    def _forward_gathered(self, fused4, dur_cum, mel_len, T, zero_padded):
        dev = fused4.device
        B, N, _ = fused4.shape
        with torch.cuda.device(dev):
            h = self._backend.ensure(dev)
            mel = torch.empty(B, T, self.n_mel_channels, dtype=torch.float32, device=dev)
            ws, wsn = self._backend.scratch(dev, B, N, T)
            _cabi.check(_cabi.load().es_decoder_forward_gathered(
                h, _stream(dev), B, N, T, fused4.data_ptr(), dur_cum.data_ptr(), mel_len.data_ptr(),
                1 if zero_padded else 0, mel.data_ptr(), ws, wsn))
        return mel


class PhonemeEncoder(nn.Module):
    """Encodes phonemes to acoustic features -- drop-in for layers/networks.py:307-401.

    ``forward(x: dict, train=False) -> dict`` with the reference's keys (``pitch``, ``energy``,
    ``duration`` [B,N,1] f32, ``mel_len`` [B] int32, ``features`` [B,T,4d], ``masks`` [B,T,4d]
    bool or None when B == 1).  Extra keys prefixed ``_`` carry the un-expanded tensors that
    ``Phoneme2Mel`` feeds to the gather-fused decoder.
    """

    def __init__(self, pitch_stats=None, energy_stats=None, depth=2, reduction=4, head=1, embed_dim=128,
                 kernel_size=3, expansion=1):
        super().__init__()
        if pitch_stats is None or energy_stats is None:
            raise ValueError("pitch_stats and energy_stats (min, max) are required (model.py:127-130)")
        self.encoder = Encoder(depth=depth, reduction=reduction, head=head, embed_dim=embed_dim,
                               kernel_size=kernel_size, expansion=expansion)
        dim = embed_dim // reduction
        self.fuse = Fuse(self.encoder.get_feature_dims(), kernel_size=kernel_size)
        self.feature_upsampler = FeatureUpsampler()
        self.pitch_decoder = AcousticDecoder(dim, pitch_stats=pitch_stats)
        self.energy_decoder = AcousticDecoder(dim, energy_stats=energy_stats)
        self.duration_decoder = AcousticDecoder(dim, duration=True)
        cfg = ESConfig(depth=depth, reduction=reduction, head=head, embed_dim=embed_dim,
                       kernel_size=kernel_size, expansion=expansion,
                       pitch_stats=tuple(float(v) for v in pitch_stats),
                       energy_stats=tuple(float(v) for v in energy_stats))
        cfg.validate()
        self._cfg = cfg
        self._backend = _Backend(self, cfg, "encoder")
        # True: forward() materialises "features"/"masks" like the reference.  Phoneme2Mel sets
        # it per call: the decoder consumes the un-expanded tensors, so inference never needs them.
        self.materialize_features = True

    def set_fused_phoneme(self, enable: bool) -> None:
        """True (default): the whole phoneme side runs as ONE kernel when the geometry allows it (tiny,
        N <= 128; include/es_b200.h: es_model_set_fused_phoneme); False: one launch per layer."""
        self._backend.fused_phoneme = bool(enable)

    def _core(self, x, train):
        phoneme = x["phoneme"]
        _require_cuda(phoneme, 'x["phoneme"]')
        dev = phoneme.device
        B, N = phoneme.shape
        d = self._cfg.dim
        ids = phoneme.to(torch.int32).contiguous()
        # networks.py:338 drops the mask when the batch holds ONE utterance.  A shard of a larger batch
        # (sharding.shard_batch) can have one row too: it carries x["global_batch_size"], and the mask decision
        # follows the GLOBAL batch so that shards stay bit-comparable with the unsharded run.
        gB = int(x.get("global_batch_size", B))
        mask = x["phoneme_mask"] if gB > 1 else None
        mask_u8 = None
        if mask is not None:
            mask_u8 = mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
        pt = et = dt = None
        if train:                                                       # networks.py:340-343
            pt = x["pitch"].to(device=dev, dtype=torch.float32).contiguous()
            et = x["energy"].to(device=dev, dtype=torch.float32).contiguous()
            dt = x["duration"].to(device=dev, dtype=torch.int32).contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        pitch = torch.empty(B, N, 1, **f32)
        energy = torch.empty(B, N, 1, **f32)
        dur = torch.empty(B, N, 1, **f32)
        fused4 = torch.empty(B, N, 4 * d, **f32)
        dur_int = torch.empty(B, N, **i32)
        dur_cum = torch.empty(B, N, **i32)
        mel_len = torch.empty(B, **i32)
        with torch.cuda.device(dev):
            h = self._backend.ensure(dev)
            ws, wsn = self._backend.scratch(dev, B, N, 0)
            _cabi.check(_cabi.load().es_encoder_forward(
                h, _stream(dev), B, N, ids.data_ptr(), _ptr(mask_u8), _ptr(pt), _ptr(et), _ptr(dt),
                pitch.data_ptr(), energy.data_ptr(), dur.data_ptr(), fused4.data_ptr(),
                dur_int.data_ptr(), dur_cum.data_ptr(), mel_len.data_ptr(), ws, wsn))
        if train:
            # networks.py:344 (torch.max(mel_len).item()); a caller that already knows the bound
            # can pass it as x["max_mel_len"] (python int) and keep the stream free of host syncs
            T = x.get("max_mel_len")
            if T is None:
                T = int(x["mel_len"].max().item())
            T = int(T)
        else:
            T = int(mel_len.max().item())                               # networks.py:246 (one sync, not B)
            # the stream is idle here anyway: poll the device-side mbarrier-timeout flag (covers this call's
            # phoneme-side kernels and the previous call's decoder), so a pipeline fault cannot pass silently
            with torch.cuda.device(dev):
                _cabi.check(_cabi.load().es_check_async_errors(_stream(dev)))
        return {"pitch": pitch, "energy": energy, "duration": dur, "mel_len": mel_len,
                "_fused4": fused4, "_dur_cum": dur_cum, "_dur_int": dur_int, "_T": T,
                "_mask_u8": mask_u8, "_global_B": gB}

    def _expand(self, y):
        """FeatureUpsampler.forward (networks.py:228-258): materialise features / masks."""
        fused4, T = y["_fused4"], y["_T"]
        B, N, C4 = fused4.shape
        dev = fused4.device
        feats = torch.empty(B, T, C4, dtype=torch.float32, device=dev)
        fmask = torch.empty(B, T, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().es_length_regulate(
                self._backend.handle, _stream(dev), B, N, T, fused4.data_ptr(), y["_dur_cum"].data_ptr(),
                _ptr(y["_mask_u8"]), feats.data_ptr(), fmask.data_ptr(), None))
        y["features"] = feats
        # the reference returns a [B,T,4d] bool tensor; this is the same values as a broadcast view
        # (None for a single-utterance batch, networks.py:391-392)
        y["masks"] = None if y["_mask_u8"] is None else fmask.view(torch.bool).unsqueeze(-1).expand(B, T, C4)
        return y

    def forward(self, x, train=False):
        y = self._core(x, train)
        if y["_T"] <= 0:
            raise RuntimeError("all durations are zero: the utterances have no frames "
                               "(the reference raises inside its decoder convolution here)")
        if self.materialize_features:
            self._expand(y)
        return y


class Phoneme2Mel(nn.Module):
    """From Phoneme Sequence to Mel Spectrogram -- drop-in for layers/networks.py:404-434."""

    def __init__(self, encoder, decoder):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        # train=True returns the reference's dict, which includes the expanded "features" /
        # "masks"; set False to skip materialising them (nothing on the mel path reads them).
        self.return_features = True
        # model.py:156 calls phoneme2mel(x, train=True) from the Lightning training loop and differentiates the result.
        # The fused inference kernels keep no tape, so in that situation -- module in train() mode, train=True, autograd
        # enabled -- forward() goes through training.forward_train: the same parameters, the reference's dict, every
        # operator and adjoint a kernel of this library.  eval() / torch.no_grad() callers keep the fused path.
        self.autograd_when_training = True

    def set_tensor_core(self, enable: bool) -> None:
        """True (default): tcgen05 kernels wherever a layer is inside their envelope; False: fp32 SIMT only."""
        self.decoder.set_tensor_core(enable)
        self.encoder._backend.tensor_core = bool(enable)

    @staticmethod
    def check_async_errors(device=None) -> None:
        """Synchronise the current stream and raise if a tcgen05 kernel reported an mbarrier timeout
        (the kernels bound every wait instead of hanging the GPU)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().es_check_async_errors(_stream(dev)))

    def capture(self, x, train=True):
        """CUDA-graph the whole forward for a fixed batch geometry (B, N, T).

        The path is ~30 small launches; replaying one graph removes the per-launch host cost.
        Needs a host-sync-free call: teacher-forced (``train=True``) with ``x["max_mel_len"]``
        given.  Returns a ``GraphedForward``: call it with a batch dict of the same shapes (its
        tensors are copied into static buffers) or with no argument to replay on the static
        inputs; the outputs live in static buffers that are overwritten by the next replay."""
        return GraphedForward(self, x, train)

    def forward(self, x, train=False):
        if isinstance(x, list):                                         # networks.py:418
            x = x[0]
        if train and self.training and self.autograd_when_training and torch.is_grad_enabled():
            from . import training
            return training.forward_train(self, x)
        pred = self.encoder._core(x, train)
        T = pred["_T"]
        if T <= 0:
            raise RuntimeError("all durations are zero: the utterances have no frames "
                               "(the reference raises inside its decoder convolution here)")
        # decoder with the length-regulator gather fused into its first kernel; padded frames are
        # zeroed only when the (global) batch has more than one utterance, like the reference (networks.py:424-427)
        mel = self.decoder._forward_gathered(pred["_fused4"], pred["_dur_cum"], pred["mel_len"], T,
                                             zero_padded=pred["_global_B"] > 1)
        pred["mel"] = mel
        if train:
            if self.return_features:
                self.encoder._expand(pred)
            return pred
        return mel, pred["mel_len"], pred["duration"]


class GraphedForward:
    """One captured CUDA graph of ``Phoneme2Mel.forward`` (see ``Phoneme2Mel.capture``)."""

    def __init__(self, model, x, train=True):
        if not train or x.get("max_mel_len") is None:
            raise ValueError("graph capture needs the sync-free call: train=True and x['max_mel_len']")
        self.model, self.train = model, train
        self.static_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in x.items()}
        dev = self.static_in["phoneme"].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # packs weights, sets kernel attributes, sizes the workspace
                model(self.static_in, train=train)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = model(self.static_in, train=train)
        # the graph holds raw pointers into the backends' workspaces and packed weight images: keep them alive here
        # and remember which weight version / switches they belong to
        self._pinned = [(b, b.key, b.workspace, b.flat) for b in (model.encoder._backend, model.decoder._backend)]

    def _check_current(self):
        for b, key, ws, flat in self._pinned:
            params = list(b.owner.parameters())
            now = key[:4] + tuple((p.data_ptr(), p._version) for p in params)
            if now != key:
                raise RuntimeError("GraphedForward: the model's weights, device or kernel switches changed after capture; "
                                   "capture again (the graph replays the packed image it was captured with)")

    def __call__(self, x=None):
        self._check_current()
        if x is not None:
            T = x.get("max_mel_len")
            if T is not None and int(T) != int(self.static_in["max_mel_len"]):
                raise ValueError(f"GraphedForward was captured with max_mel_len={self.static_in['max_mel_len']}, got {T}")
            for k, v in x.items():
                if torch.is_tensor(v):
                    if tuple(v.shape) != tuple(self.static_in[k].shape):
                        raise ValueError(f"GraphedForward: x[{k!r}] has shape {tuple(v.shape)}, captured {tuple(self.static_in[k].shape)}")
                    self.static_in[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.static_out
