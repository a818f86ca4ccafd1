"""GPU parity tests proper: the CUDA path (through the C ABI) vs the reference fixtures and the oracle.

Bars (BASELINE.json north_star / SURVEY.md section 8c):
  mel max-abs <= 1e-3 vs the reference fp32 CPU path; scalar predictions <= 1e-4;
  every integer (mel_len, rounded durations, bucket indices via the embedded rows,
  length-regulator source indices, masks) bit-exact.
"""
import os

import numpy as np
import pytest
import torch

import efficientspeech_b200 as es
from efficientspeech_b200 import _cabi
from efficientspeech_b200.config import VARIANTS
from efficientspeech_b200.params import init_state_dict
from efficientspeech_b200.synthetic import make_batch
from helpers import TOL_MEL, TOL_PRED, golden_files, load_golden
from oracle import es_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cuda_model(vname, sd):
    m = es.build_model(vname)
    es.load_numpy_state(m, sd)
    return m.to(DEV).eval()


def to_dev(batch):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(DEV) for k, v in batch.items()}


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_reference_fixture_parity(path):
    vname, cfg, sd, batch, g = load_golden(path)
    model = cuda_model(vname, sd)
    x = to_dev(batch)
    with torch.no_grad():
        tf = model(x, train=True)
        mel_fr, len_fr, dur_fr = model(x, train=False)
    assert tf["mel_len"].dtype == torch.int32
    assert np.array_equal(npy(tf["mel_len"]), g["tf_mel_len"])
    assert tuple(tf["mel"].shape) == g["tf_mel"].shape
    assert np.abs(npy(tf["mel"]) - g["tf_mel"]).max() <= TOL_MEL
    for k in ("pitch", "energy", "duration"):
        assert tuple(tf[k].shape) == g["tf_" + k].shape
        assert np.abs(npy(tf[k]) - g["tf_" + k]).max() <= TOL_PRED, k
    # free running: durations are rounded predictions -> same T, same lengths, same mel
    assert np.array_equal(npy(len_fr), g["fr_mel_len"])
    assert tuple(mel_fr.shape) == g["fr_mel"].shape
    assert np.abs(npy(mel_fr) - g["fr_mel"]).max() <= TOL_MEL
    assert np.abs(npy(dur_fr) - g["fr_duration"]).max() <= TOL_PRED
    # length-regulator source map from the reference FeatureUpsampler: bit-exact
    B, N = batch["phoneme"].shape
    T = int(g["tf_mel_len"].max())
    src = torch.empty(B, T, dtype=torch.int32, device=DEV)
    _cabi.check(_cabi.load().es_length_regulate(
        model.encoder._backend.handle, torch.cuda.current_stream().cuda_stream, B, N, T, None,
        tf["_dur_cum"].data_ptr(), None, None, None, src.data_ptr()))
    assert np.array_equal(npy(src), g["tf_src"])


CASES = [("tiny", 5, 40, True), ("tiny", 1, 19, False), ("tiny", 3, 129, True), ("tiny", 2, 2, False),
         ("small", 4, 33, True), ("small", 1, 64, False), ("base", 3, 31, True), ("base", 2, 70, True),
         ("tiny", 7, 300, True)]


@pytest.mark.parametrize("vname,B,N,ragged", CASES, ids=lambda v: str(v))
def test_oracle_parity_all_outputs(vname, B, N, ragged):
    cfg = VARIANTS[vname]
    sd = init_state_dict(cfg, seed=100 + B + N)
    batch = make_batch(cfg, B, N, seed=B * 7 + N, ragged=ragged, fixed_duration=None, max_dur=6)
    model = cuda_model(vname, sd)
    x = to_dev(batch)
    for train in (True, False):
        o = es_oracle.phoneme2mel(batch, sd, train=train)
        with torch.no_grad():
            out = model(x, train=True) if train else model.encoder(x, train=False)
            mel = out["mel"] if train else model(x, train=False)[0]
        assert np.array_equal(npy(out["mel_len"]), o["mel_len"])
        assert np.array_equal(npy(out["_dur_int"]), o["_dur_int"])                   # rounded durations: exact
        assert np.abs(npy(out["_fused4"]) - o["_fused4"]).max() <= TOL_PRED * 5        # a bucket flip would be O(1)
        assert np.abs(npy(out["features"]) - o["features"]).max() <= TOL_PRED * 5
        if B > 1:
            assert np.array_equal(npy(out["masks"]), o["masks"])                     # bool [B,T,4d]: exact
        else:
            assert out["masks"] is None and o["masks"] is None
        for k in ("pitch", "energy", "duration"):
            assert np.abs(npy(out[k]) - o[k]).max() <= TOL_PRED, k
        assert tuple(mel.shape) == o["mel"].shape
        assert np.abs(npy(mel) - o["mel"]).max() <= TOL_MEL


@pytest.mark.parametrize("vname", list(VARIANTS))
def test_mel_decoder_standalone(vname):
    cfg = VARIANTS[vname]
    sd = init_state_dict(cfg, seed=5)
    model = cuda_model(vname, sd)
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((3, 77, cfg.dx4)).astype(np.float32)
    S = es_oracle._cast_state(sd, np.float32)
    want = es_oracle.mel_decoder(feats, S, es_oracle.infer_config(S))
    with torch.no_grad():
        got = model.decoder(torch.from_numpy(feats).to(DEV))
    assert np.abs(npy(got) - want).max() <= TOL_MEL


def test_length_regulator_exact_random():
    model = cuda_model("tiny", init_state_dict(VARIANTS["tiny"], seed=1))
    model.encoder._backend.ensure(torch.device(DEV))
    lib = _cabi.load()
    rng = np.random.default_rng(123)
    for trial in range(25):
        B, N = int(rng.integers(1, 9)), int(rng.integers(1, 200))
        dur = rng.integers(0, 9, size=(B, N)).astype(np.int32)
        if trial % 5 == 0:
            dur[rng.integers(0, B)] = 0                                      # an utterance with no frames
        if trial % 7 == 0:
            dur[:, : N // 2] = 0                                             # leading zero-duration phonemes
        feats = rng.standard_normal((B, N, 128)).astype(np.float32)
        pmask = rng.random((B, N)) < 0.1
        f, m, ml, src = es_oracle.feature_upsampler(feats, np.repeat(pmask[..., None], 128, 2), dur)
        T = int(ml.max())
        if T == 0:
            continue
        cum = torch.from_numpy(np.cumsum(dur, axis=1).astype(np.int32)).to(DEV)
        tf = torch.from_numpy(feats).to(DEV)
        tm = torch.from_numpy(pmask).to(DEV).view(torch.uint8)
        out = torch.empty(B, T, 128, device=DEV)
        fm = torch.empty(B, T, dtype=torch.uint8, device=DEV)
        sr = torch.empty(B, T, dtype=torch.int32, device=DEV)
        _cabi.check(lib.es_length_regulate(model.encoder._backend.handle, torch.cuda.current_stream().cuda_stream,
                                           B, N, T, tf.data_ptr(), cum.data_ptr(), tm.data_ptr(),
                                           out.data_ptr(), fm.data_ptr(), sr.data_ptr()))
        assert np.array_equal(npy(sr), src)
        assert np.array_equal(npy(out), f)                                   # pure copy: bit-exact
        assert np.array_equal(npy(fm).astype(bool), m[..., 0])


def test_full_size_properties_tiny_b256():
    """BASELINE configs[1] shape: size-independent properties + spot parity vs the oracle."""
    cfg = VARIANTS["tiny"]
    sd = init_state_dict(cfg, seed=0)
    model = cuda_model("tiny", sd)
    B, N = 256, 128
    batch = make_batch(cfg, B, N, seed=0, ragged=False, fixed_duration=6)
    x = to_dev(batch)
    x["max_mel_len"] = 6 * N
    with torch.no_grad():
        out = model(x, train=True)
    mel = npy(out["mel"])
    assert mel.shape == (B, 6 * N, 80) and np.isfinite(mel).all()
    assert np.array_equal(npy(out["mel_len"]), batch["mel_len"])
    # utterances are independent: a shard of the batch gives bit-identical results
    idx = [0, 100, 255]
    sub = {k: v[idx] for k, v in batch.items()}
    xs = to_dev(sub)
    xs["max_mel_len"] = 6 * N
    with torch.no_grad():
        outs = model(xs, train=True)
    assert np.array_equal(npy(outs["mel"]), mel[idx])
    o = es_oracle.phoneme2mel(sub, sd, train=True)
    assert np.abs(mel[idx] - o["mel"]).max() <= TOL_MEL
    # ragged: padded frames are exactly zero, lengths are the duration sums
    rb = make_batch(cfg, 64, N, seed=3, ragged=True, fixed_duration=None)
    with torch.no_grad():
        ro = model(to_dev(rb), train=True)
    rmel, rlen = npy(ro["mel"]), npy(ro["mel_len"])
    assert np.array_equal(rlen, rb["duration"].sum(1))
    for b in range(64):
        assert (rmel[b, rlen[b]:] == 0).all()


@pytest.mark.parametrize("B,N,ragged", [(5, 40, True), (2, 64, False), (3, 129, True)])
def test_tensor_core_decoder_matches_simt_and_oracle(B, N, ragged):
    """tcgen05 split-fp16 decoder (default) vs the fp32 SIMT decoder vs the oracle; ragged T exercises
    partial tiles, utterance-boundary halos and zero-length tails."""
    cfg = VARIANTS["tiny"]
    sd = init_state_dict(cfg, seed=77)
    batch = make_batch(cfg, B, N, seed=N, ragged=ragged, fixed_duration=None, max_dur=9)
    model = cuda_model("tiny", sd)
    x = to_dev(batch)
    with torch.no_grad():
        model.set_tensor_core(True)
        tc = npy(model(x, train=True)["mel"])
        model.check_async_errors()
        model.set_tensor_core(False)
        simt = npy(model(x, train=True)["mel"])
    o = es_oracle.phoneme2mel(batch, sd, train=True)
    assert np.abs(simt - o["mel"]).max() <= TOL_MEL
    assert np.abs(tc - o["mel"]).max() <= TOL_MEL
    assert np.abs(tc - simt).max() <= 2e-4        # both are fp32-class: far inside the 1e-3 bar
    # standalone MelDecoder.forward through the tensor-core path (plain projection prologue)
    model.set_tensor_core(True)
    feats = torch.from_numpy(o["features"]).to(DEV)
    S = es_oracle._cast_state(sd, np.float32)
    want = es_oracle.mel_decoder(o["features"], S, es_oracle.infer_config(S))
    with torch.no_grad():
        got = npy(model.decoder(feats))
    model.check_async_errors()
    assert np.abs(got - want).max() <= TOL_MEL


def test_errors_are_python_exceptions():
    model = cuda_model("tiny", init_state_dict(VARIANTS["tiny"], seed=1))
    with pytest.raises(RuntimeError):
        model.decoder(torch.zeros(1, 8, 64, device=DEV))                     # wrong channel count
    cfg = VARIANTS["tiny"]
    b = make_batch(cfg, 2, 8, seed=0)
    b["duration"][:] = 0
    b["mel_len"][:] = 0
    with pytest.raises(RuntimeError, match="no frames"):
        model(to_dev(b), train=True)


def test_cuda_graph_replay_matches_eager():
    cfg = VARIANTS["tiny"]
    sd = init_state_dict(cfg, seed=21)
    model = cuda_model("tiny", sd)
    batch = make_batch(cfg, 6, 48, seed=5, ragged=True, fixed_duration=None, max_dur=7)
    x = to_dev(batch)
    x["max_mel_len"] = int(batch["mel_len"].max())
    with torch.no_grad():
        eager = npy(model(x, train=True)["mel"])
    g = model.capture(x, train=True)
    assert np.array_equal(npy(g()["mel"]), eager)
    # new inputs of the same geometry through the same graph
    batch2 = make_batch(cfg, 6, 48, seed=6, ragged=True, fixed_duration=None, max_dur=7)
    batch2["duration"] = np.minimum(batch2["duration"], 3)            # stays within the captured T
    batch2["mel_len"] = batch2["duration"].sum(1).astype(np.int32)
    x2 = to_dev(batch2)
    x2["max_mel_len"] = x["max_mel_len"]
    with torch.no_grad():
        want = npy(model(x2, train=True)["mel"])
    got = npy(g({k: v for k, v in x2.items() if torch.is_tensor(v)})["mel"])
    assert np.array_equal(got, want)
    model.check_async_errors()


def test_frame_rows_exact_random():
    """The frame -> table-row map of the gathered decoder entry (es_gather.cu) against the oracle's
    FeatureUpsampler source indices: bit-exact, including zero-duration phonemes, empty utterances and
    N beyond the shared-memory staging limit."""
    model = cuda_model("tiny", init_state_dict(VARIANTS["tiny"], seed=1))
    model.decoder._backend.ensure(torch.device(DEV))
    lib = _cabi.load()
    rng = np.random.default_rng(321)
    shapes = [(int(rng.integers(1, 9)), int(rng.integers(1, 300))) for _ in range(20)] + [(2, 9000)]
    for trial, (B, N) in enumerate(shapes):
        dur = rng.integers(0, 9, size=(B, N)).astype(np.int32)
        if trial % 5 == 0:
            dur[rng.integers(0, B)] = 0
        if trial % 7 == 0:
            dur[:, : N // 2] = 0
        feats = np.zeros((B, N, 4), np.float32)
        _, _, ml, src = es_oracle.feature_upsampler(feats, np.zeros((B, N, 4), bool), dur)
        T = int(ml.max())
        if T == 0:
            continue
        want = np.where(src >= 0, src + np.arange(B)[:, None] * N, B * N).astype(np.int32)
        cum = torch.from_numpy(np.cumsum(dur, axis=1).astype(np.int32)).to(DEV)
        mlen = torch.from_numpy(ml.astype(np.int32)).to(DEV)
        rows = torch.empty(B, T, dtype=torch.int32, device=DEV)
        _cabi.check(lib.es_frame_rows(model.decoder._backend.handle, torch.cuda.current_stream().cuda_stream,
                                      B, N, T, cum.data_ptr(), mlen.data_ptr(), rows.data_ptr()))
        assert np.array_equal(npy(rows), want)


@pytest.mark.parametrize("vname,B,N,max_dur", [("tiny", 5, 40, 9), ("tiny", 3, 129, 4), ("tiny", 1, 33, 7),
                                               ("tiny", 4, 64, 40), ("small", 4, 33, 6), ("base", 2, 31, 6)])
def test_gather_modes_agree(vname, B, N, max_dur):
    """The two ways of joining the length regulator and the decoder (ES_GATHER_*): projection per phoneme + row-gather
    kernel, projection per phoneme + gathered loads in the first block.  Both within the mel bar of the oracle and
    bit-identical to each other."""
    cfg = VARIANTS[vname]
    sd = init_state_dict(cfg, seed=31 + N)
    batch = make_batch(cfg, B, N, seed=N + B, ragged=B > 1, fixed_duration=None, max_dur=max_dur)
    model = cuda_model(vname, sd)
    x = to_dev(batch)
    o = es_oracle.phoneme2mel(batch, sd, train=True)
    got = {}
    try:
        for mode in (1, 2):
            model.decoder.set_gather_mode(mode)
            with torch.no_grad():
                got[mode] = npy(model(x, train=True)["mel"])
            model.check_async_errors()
            assert np.abs(got[mode] - o["mel"]).max() <= TOL_MEL, mode
    finally:
        model.decoder.set_gather_mode(2)
    assert np.array_equal(got[1], got[2])


def test_gather_fused_zero_duration_run():
    """A 75-phoneme run of zero durations in the middle of an utterance: the distinct table rows of one tile
    no longer fit the 68-row slot and the gathered layer falls back to one copy per frame (es_umma_dec.cu)."""
    cfg = VARIANTS["tiny"]
    sd = init_state_dict(cfg, seed=12)
    batch = make_batch(cfg, 2, 200, seed=4, ragged=False, fixed_duration=None, max_dur=5)
    batch["duration"][:, 20:95] = 0
    batch["duration"][:, 19] = batch["duration"][:, 95] = 3
    batch["mel_len"] = batch["duration"].sum(1).astype(np.int32)
    model = cuda_model("tiny", sd)
    x = to_dev(batch)
    o = es_oracle.phoneme2mel(batch, sd, train=True)
    got = {}
    try:
        for mode in (1, 2):
            model.decoder.set_gather_mode(mode)
            with torch.no_grad():
                got[mode] = npy(model(x, train=True)["mel"])
            model.check_async_errors()
    finally:
        model.decoder.set_gather_mode(2)
    assert np.abs(got[2] - o["mel"]).max() <= TOL_MEL
    assert np.array_equal(got[1], got[2])


@pytest.mark.parametrize("B,N,ragged,train", [(5, 40, True, True), (3, 128, True, False), (1, 19, False, True),
                                              (150, 50, True, True), (2, 2, False, False), (4, 127, True, True),
                                              (300, 9, True, False)])
def test_fused_phoneme_kernel_matches_per_layer(B, N, ragged, train):
    """The one-kernel phoneme side (es_umma_phoneme.cu, tiny geometry, N <= 128) against the per-layer launches:
    integers identical, floats within fp32 rounding of each other, and the end-to-end mel within the bar of the
    oracle.  B > 148 makes CTAs loop over several utterances; odd N exercises the ragged half-rate level."""
    cfg = VARIANTS["tiny"]
    sd = init_state_dict(cfg, seed=40 + N)
    batch = make_batch(cfg, B, N, seed=B + 3 * N, ragged=ragged, fixed_duration=None, max_dur=5)
    model = cuda_model("tiny", sd)
    x = to_dev(batch)
    lib = _cabi.load()
    out = {}
    try:
        for fused in (True, False):
            model.encoder.set_fused_phoneme(fused)
            l0 = lib.es_launch_count()
            with torch.no_grad():
                y = model.encoder(x, train=train)
            out[fused] = {k: npy(y[k]) for k in ("pitch", "energy", "duration", "mel_len", "_fused4", "_dur_int", "_dur_cum")}
            out[fused]["launches"] = lib.es_launch_count() - l0
            model.check_async_errors()
    finally:
        model.encoder.set_fused_phoneme(True)
    assert out[True]["launches"] < out[False]["launches"] and out[True]["launches"] <= 2      # fused kernel (+ length regulator)
    for k in ("mel_len", "_dur_int", "_dur_cum"):
        assert np.array_equal(out[True][k], out[False][k]), k
    for k in ("pitch", "energy", "duration"):
        assert np.abs(out[True][k] - out[False][k]).max() <= 2e-5, k
    assert np.abs(out[True]["_fused4"] - out[False]["_fused4"]).max() <= 5e-5
    if B <= 8:
        o = es_oracle.phoneme2mel(batch, sd, train=train)
        with torch.no_grad():
            mel = model(x, train=True)["mel"] if train else model(x, train=False)[0]
        assert np.abs(npy(mel) - o["mel"]).max() <= TOL_MEL


@pytest.mark.parametrize("B,train", [(1, True), (7, True), (64, False), (300, True), (2000, True)])
def test_collate_on_device_exact(B, train):
    """es_collate (datamodule.py:29-76 on the device) against the oracle: permutation (stable), padded arrays, mask,
    lengths and mel_len bit-exact, including equal lengths and zero durations."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_collate_cpu import ragged_items
    items = ragged_items(1000 + B, B, lo=1, hi=128 if B < 500 else 24, train=train)
    got = es.collate(items, DEV)
    o = es_oracle.collate(items)
    keys = ["perm", "phoneme", "phoneme_len", "phoneme_mask"] + (["pitch", "energy", "duration", "mel_len"] if train else [])
    assert set(k for k in got) == set(keys)
    for k in keys:
        g = npy(got[k])
        assert g.dtype == o[k].dtype or k == "phoneme_mask", (k, g.dtype, o[k].dtype)
        assert np.array_equal(g, o[k]), k
    assert got["phoneme_mask"].dtype == torch.bool
    if train and B > 1:
        # the collated batch feeds the model directly
        cfg = VARIANTS["tiny"]
        model = cuda_model("tiny", init_state_dict(cfg, seed=3))
        ids = torch.clamp(got["phoneme"], max=cfg.n_symbols - 1)
        x = {"phoneme": ids[:8], "phoneme_mask": got["phoneme_mask"][:8], "pitch": got["pitch"][:8], "energy": got["energy"][:8],
             "duration": got["duration"][:8], "mel_len": got["mel_len"][:8]}
        if int(x["mel_len"].max()) > 0 and x["phoneme"].shape[1] >= 2:
            with torch.no_grad():
                out = model(x, train=True)
            assert torch.isfinite(out["mel"]).all()


@pytest.mark.parametrize("vname,B,N", [("tiny", 12, 128), ("small", 6, 96), ("base", 5, 64), ("tiny", 3, 20)])
@pytest.mark.parametrize("train", [True, False])
def test_ragged_schedule_is_bit_identical(vname, B, N, train):
    """Ragged scheduling (the decoder walks only the tiles that can reach a valid frame, es_gather.cu: tile_list_kernel)
    against the dense schedule the reference implies: the same bits on every frame, zeros past mel_len -- including an
    utterance with NO frames and one that is a small fraction of the longest."""
    cfg = VARIANTS[vname]
    sd = dict(init_state_dict(cfg, seed=50 + N))
    sd["encoder.duration_decoder.linear.bias"] = np.full_like(sd["encoder.duration_decoder.linear.bias"], 5.0)
    batch = make_batch(cfg, B, N, seed=9 + B, ragged=True, fixed_duration=None, max_dur=9, min_len=4)
    batch["duration"][-1] = 0                                       # an utterance without frames
    batch["duration"][-2, 6:] = 0                                   # a very short one
    batch["mel_len"] = batch["duration"].sum(1).astype(np.int32)
    model = cuda_model(vname, sd)
    x = to_dev(batch)
    got = {}
    try:
        for on in (True, False):
            model.decoder.set_ragged_schedule(on)
            with torch.no_grad():
                out = model(x, train=True) if train else model(x, train=False)
            mel, mel_len = (out["mel"], out["mel_len"]) if train else (out[0], out[1])
            got[on] = (npy(mel), npy(mel_len))
            model.check_async_errors()
    finally:
        model.decoder.set_ragged_schedule(True)
    assert np.array_equal(got[True][1], got[False][1])
    assert np.array_equal(got[True][0], got[False][0])
    ml = got[True][1]
    for b in range(B):
        assert (got[True][0][b, ml[b]:] == 0).all()
    if train:
        o = es_oracle.phoneme2mel(batch, sd, train=True)
        assert np.abs(got[True][0] - o["mel"]).max() <= TOL_MEL
