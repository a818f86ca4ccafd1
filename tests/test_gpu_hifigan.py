"""HiFi-GAN row, GPU side: es_hifigan_forward (through efficientspeech_b200.hifigan.Generator and the C ABI) against the
numpy oracle and the reference fixture.  Waveform max-abs <= 2e-4 (fp32 both sides; the kernels only reorder sums)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import efficientspeech_b200 as es
from oracle import hifigan_oracle as ho
from oracle import ref_shim

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_WAV = 2e-4


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def plain_generator(cfg, sd):
    G = es.hifigan.Generator(es.hifigan.AttrDict(cfg)).eval()
    G.remove_weight_norm()
    G.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return G.to(DEV)


def test_reference_fixture_parity():
    mg = _mg()
    z = np.load(os.path.join(ROOT, "tests", "golden", "hifigan_v2_t6.npz"))
    G = plain_generator(mg.HIFIGAN_V2, mg.hifigan_seeded_state(mg.HIFIGAN_V2, int(z["weight_seed"])))
    with torch.no_grad():
        wav = G(torch.from_numpy(z["mel"]).to(DEV))
    assert tuple(wav.shape) == z["wav_seeded"].shape
    assert np.abs(wav.cpu().numpy() - z["wav_seeded"]).max() <= TOL_WAV


@pytest.mark.parametrize("B,T", [(1, 1), (3, 37), (2, 130), (1, 768)])
def test_oracle_parity_seeded_weights(B, T):
    """Odd lengths (partial tiles at every stage), single frame, and the BASELINE utterance length T = 768."""
    mg = _mg()
    sd = mg.hifigan_seeded_state(mg.HIFIGAN_V2, 3 + T)
    G = plain_generator(mg.HIFIGAN_V2, sd)
    mel = (np.random.default_rng(T).standard_normal((B, 80, T)) * 1.5 - 4).astype(np.float32)
    want = ho.generator(mel, sd, mg.HIFIGAN_V2)
    with torch.no_grad():
        got = G(torch.from_numpy(mel).to(DEV)).cpu().numpy()
    assert got.shape == (B, 1, 256 * T)
    assert np.abs(got - want).max() <= TOL_WAV


@pytest.mark.skipif(not os.path.isfile(os.path.join(ref_shim.REF_DIR, "hifigan", "LJ_V2", "generator_v2")),
                    reason="reference checkpoint not staged (oracle/build_ref.py)")
def test_checkpoint_through_get_hifigan_flow_and_transposed_mel():
    """model.py:23-48 flow with the checkpoint of the tree: load with weight norm attached, eval, remove_weight_norm;
    the mel arrives as the transposed VIEW of the acoustic model's [B,T,80] output (model.py:160) and is consumed in place."""
    cfg, ck = ho.load_reference_checkpoint(ref_shim.REF_DIR)
    G = es.hifigan.Generator(es.hifigan.AttrDict(cfg))
    G.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()}, strict=True)
    G.eval()
    mel_btc = (np.random.default_rng(9).standard_normal((2, 50, 80)) * 1.5 - 4).astype(np.float32)
    want = ho.generator(np.ascontiguousarray(mel_btc.transpose(0, 2, 1)), ck, cfg)
    G = G.to(DEV)
    x = torch.from_numpy(mel_btc).to(DEV).transpose(1, 2)
    assert not x.is_contiguous()
    with torch.no_grad():
        before = G(x).cpu().numpy()                 # weight norm still attached: effective weights computed at pack time
        G.remove_weight_norm()
        after = G(x).cpu().numpy()
    assert np.abs(before - want).max() <= TOL_WAV
    assert np.abs(after - want).max() <= TOL_WAV
    assert np.abs(after).max() < 1.0


def test_text_to_waveform_pipeline_matches_the_chained_oracles():
    """es.synthesize: text -> ids -> device collation -> Phoneme2Mel (free running) -> HiFi-GAN, against the numpy oracles
    chained on the host (es_oracle.collate / phoneme2mel, hifigan_oracle.generator).  Rows come back in the caller's order."""
    import tempfile
    from efficientspeech_b200 import text as T
    from efficientspeech_b200.config import VARIANTS
    from efficientspeech_b200.params import init_state_dict
    from oracle import es_oracle
    mg = _mg()
    cfg = VARIANTS["tiny"]
    sd = dict(init_state_dict(cfg, seed=2))
    sd["encoder.duration_decoder.linear.bias"] = np.full_like(sd["encoder.duration_decoder.linear.bias"], 3.0)
    model = es.build_model("tiny")
    es.load_numpy_state(model, sd)
    model = model.to(DEV).eval()
    hsd = mg.hifigan_seeded_state(mg.HIFIGAN_V2, 21)
    G = plain_generator(mg.HIFIGAN_V2, hsd)
    lex_txt = "hello HH AH0 L OW1\nworld W ER1 L D\nthe DH AH0\nquick K W IH1 K\nbrown B R AW1 N\nfox F AA1 K S\n"
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(lex_txt)
    lex = T.read_lexicon(f.name)
    pc = {"preprocessing": {"text": {"language": "en", "text_cleaners": ["english_cleaners"]}}}
    texts = ["hello world", "the quick, brown fox - hello world!", "fox"]
    out = es.synthesize(texts, lex, None, model, G, pc, DEV)
    # host-side chain
    items = [{"phoneme": T.text2phoneme(lex, None, t, pc)} for t in texts]
    ob = es_oracle.collate(items)
    o = es_oracle.phoneme2mel({"phoneme": ob["phoneme"], "phoneme_mask": ob["phoneme_mask"]}, sd, train=False)
    wav = ho.generator(np.ascontiguousarray(o["mel"].transpose(0, 2, 1)), hsd, mg.HIFIGAN_V2)[:, 0]
    inv = np.argsort(ob["perm"])
    assert np.array_equal(out["mel_len"].cpu().numpy(), o["mel_len"][inv])
    assert np.array_equal(out["wav_len"].cpu().numpy(), o["mel_len"][inv] * 256)
    assert np.abs(out["mel"].cpu().numpy() - o["mel"][inv]).max() <= 1e-3
    assert np.abs(out["wav"].cpu().numpy() - wav[inv]).max() <= 5e-4
