"""Text front-end (row f4), CPU: symbol table, text_to_sequence and text2phoneme against the reference's own sources
(text/ package staged under oracle/_ref, synthesize.py functions compiled from the mounted tree when present)."""
import os
import sys
import tempfile

import numpy as np
import pytest

from efficientspeech_b200 import text as T
from oracle import ref_shim

CFG = {"preprocessing": {"text": {"language": "en", "text_cleaners": ["english_cleaners"]}}}
LEX = """hello HH AH0 L OW1
HELLO HH EH0 L OW1
world W ER1 L D
the DH AH0
quick K W IH1 K
brown B R AW1 N
fox F AA1 K S
it's IH1 T S
"""
SENTENCES = ["Hello world.", "the quick, brown fox!", "it's the world; hello - fox?", "hello  world"]


def _lexicon():
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(LEX)
    return f.name


def test_symbol_table_shape_and_round_trip():
    assert len(T.symbols) == 152 and len(set(T.symbols)) == 152          # + the padding id 0 -> n_symbols = 153 (networks.py:32)
    assert T.symbols[0] == "_" and T.symbols[-3:] == ["@sp", "@spn", "@sil"]
    seq = T.text_to_sequence("{HH AH0 L OW1 sp W ER1 L D}", ["english_cleaners"])
    assert len(seq) == 9 and T.sequence_to_text(seq) == "{HH AH0 L OW1 sp W ER1 L D}"


def test_text2phoneme_without_g2p_raises_on_oov_only():
    lex = T.read_lexicon(_lexicon())
    assert lex["hello"] == ["HH", "AH0", "L", "OW1"]                      # first pronunciation wins (synthesize.py:26-27)
    ids = T.text2phoneme(lex, None, "Hello, world.", CFG)                  # punctuation -> sp; spaces vanish; trailing "." stripped
    assert T.sequence_to_text(ids) == "{HH AH0 L OW1 sp W ER1 L D}"
    with pytest.raises(KeyError):
        T.text2phoneme(lex, None, "hello zyzzyva", CFG)


@pytest.mark.skipif(not os.path.isfile(os.path.join(ref_shim.REF_DIR, "text", "symbols.py")), reason="reference text/ not staged")
def test_symbols_and_text_to_sequence_match_reference_package():
    ref_shim.import_reference_layers()                                     # installs the unidecode / inflect stubs, sys.path
    import text as ref_text
    from text.symbols import symbols as ref_symbols
    assert T.symbols == list(ref_symbols)
    for s in ("{HH AH0 L OW1 sp W ER1 L D}", "{sp}", "{DH AH0} {K W IH1 K}", "{AA1 spn ZH sil}"):
        assert T.text_to_sequence(s, ["english_cleaners"]) == ref_text.text_to_sequence(s, ["english_cleaners"]), s


@pytest.mark.skipif(not os.path.isfile(os.path.join(ref_shim._MOUNTED, "synthesize.py")), reason="reference tree not mounted")
def test_text2phoneme_matches_reference_source():
    ref_shim.import_reference_layers()
    import text as ref_text
    fns, ns = ref_shim.reference_functions("synthesize.py", ["read_lexicon", "text2phoneme"])
    import re
    from string import punctuation
    ns.update({"re": re, "punctuation": punctuation, "text_to_sequence": ref_text.text_to_sequence})
    g2p = lambda w: list(w.upper()) if w.strip() else []                   # a stand-in for g2p_en.G2p, same on both sides
    path = _lexicon()
    ref_lex, lex = fns["read_lexicon"](path), T.read_lexicon(path)
    assert ref_lex == lex
    for s in SENTENCES + ["hello zebra world"]:
        want = fns["text2phoneme"](ref_lex, g2p, s, CFG)
        got = T.text2phoneme(lex, g2p, s, CFG)
        assert np.array_equal(got, want), s
