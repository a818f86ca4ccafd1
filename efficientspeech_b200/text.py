"""Text front-end of the reference, the step before collation: lexicon lookup, the
phoneme-string convention and the symbol table that turn a sentence into the int phoneme ids the acoustic model takes.

Mirrors ``synthesize.py:19-64`` (``read_lexicon``, ``text2phoneme``) and ``text/__init__.py`` + ``text/symbols.py``
(``text_to_sequence``, 152 symbols: pad, "-/", punctuation, letters, 84 ARPAbet phones prefixed "@", 3 silences).
CPU string work -- nothing here runs on the GPU, by design (the reference's own comment on this row: lowest value on
B200).  Out-of-vocabulary words go to the caller's ``g2p`` callable exactly as upstream (``g2p_en.G2p`` there; that
package is not in this image, so ``g2p=None`` raises on an OOV word instead of guessing).  ``text2phoneme`` always
emits a fully braced ARPAbet string, so the cleaners of ``text_to_sequence`` only ever see the empty text between
braces; ``english_cleaners`` is implemented as lower-casing + whitespace collapsing (number / abbreviation expansion
and transliteration need inflect / unidecode and only matter for raw text outside braces).
"""
from __future__ import annotations

import re
from string import punctuation
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

__all__ = ["symbols", "text_to_sequence", "sequence_to_text", "read_lexicon", "text2phoneme"]

# the CMU phone set with stress markers (text/cmudict.py:6-91 lists the same 84 symbols)
_VOWELS = ["AA", "AE", "AH", "AO", "AW", "AY", "EH", "ER", "EY", "IH", "IY", "OW", "OY", "UH", "UW"]
_CONSONANTS = ["B", "CH", "D", "DH", "F", "G", "HH", "JH", "K", "L", "M", "N", "NG", "P", "R", "S", "SH", "T", "TH", "V",
               "W", "Y", "Z", "ZH"]
valid_arpabet = sorted([v + s for v in _VOWELS for s in ("", "0", "1", "2")] + _CONSONANTS)

_pad = "_"
_punctuation = "!'(),.:;? "
_special = "-/"
_letters = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
_silences = ["@sp", "@spn", "@sil"]
# text/symbols.py:22-31
symbols: List[str] = [_pad] + list(_special) + list(_punctuation) + list(_letters) + ["@" + s for s in valid_arpabet] + _silences
_symbol_to_id = {s: i for i, s in enumerate(symbols)}
_id_to_symbol = {i: s for i, s in enumerate(symbols)}
_curly_re = re.compile(r"(.*?)\{(.+?)\}(.*)")
_whitespace_re = re.compile(r"\s+")


def _basic_cleaners(text: str) -> str:
    return re.sub(_whitespace_re, " ", text.lower())


_CLEANERS = {"basic_cleaners": _basic_cleaners, "english_cleaners": _basic_cleaners, "transliteration_cleaners": _basic_cleaners}


def _clean_text(text: str, cleaner_names: Sequence[str]) -> str:
    for name in cleaner_names:
        if name not in _CLEANERS:
            raise Exception("Unknown cleaner: %s" % name)
        text = _CLEANERS[name](text)
    return text


def _keep(s: str) -> bool:
    return s in _symbol_to_id and s != "_" and s != "~"


def _symbols_to_sequence(syms) -> List[int]:
    return [_symbol_to_id[s] for s in syms if _keep(s)]


def text_to_sequence(text: str, cleaner_names: Sequence[str]) -> List[int]:
    """text/__init__.py:15-44: ids of the symbols of ``text``; ``{...}`` encloses ARPAbet."""
    sequence: List[int] = []
    while len(text):
        m = _curly_re.match(text)
        if not m:
            sequence += _symbols_to_sequence(_clean_text(text, cleaner_names))
            break
        sequence += _symbols_to_sequence(_clean_text(m.group(1), cleaner_names))
        sequence += _symbols_to_sequence(["@" + s for s in m.group(2).split()])
        text = m.group(3)
    return sequence


def sequence_to_text(sequence) -> str:
    """text/__init__.py:47-58."""
    out = ""
    for i in sequence:
        s = _id_to_symbol.get(int(i))
        if s is None:
            continue
        out += "{%s}" % s[1:] if len(s) > 1 and s[0] == "@" else s
    return out.replace("}{", " ")


def read_lexicon(lex_path: str) -> Dict[str, List[str]]:
    """synthesize.py:19-28: first pronunciation of every (lower-cased) word."""
    lexicon: Dict[str, List[str]] = {}
    with open(lex_path) as f:
        for line in f:
            parts = re.split(r"\s+", line.strip("\n"))
            word = parts[0].lower()
            if word not in lexicon:
                lexicon[word] = parts[1:]
    return lexicon


def text2phoneme(lexicon: Dict[str, List[str]], g2p: Optional[Callable[[str], List[str]]], text: str, preprocess_config: dict,
                 verbose: bool = False) -> np.ndarray:
    """synthesize.py:37-64: sentence -> numpy array of phoneme ids."""
    text = text.rstrip(punctuation)
    lang = preprocess_config["preprocessing"]["text"]["language"]
    phones: List[str] = []
    for w in re.split(r"([,;.\-\?\!\s+])", text):
        if w.lower() in lexicon:
            phones += lexicon[w.lower()]
        elif lang == "t1":
            phones += list(w.lower())
        else:
            if g2p is None:
                if w.strip() == "" or all(c in punctuation or c.isspace() for c in w):
                    phones += [p for p in w if p != " "]          # what G2p returns for bare punctuation / nothing
                    continue
                raise KeyError(f"{w!r} is not in the lexicon and no g2p callable was given (g2p_en is not available here)")
            phones += [p for p in g2p(w) if p != " "]
    s = "{" + "}{".join(phones) + "}"
    s = re.sub(r"\{[^\w\s]?\}", "{sp}", s)
    s = s.replace("}{", " ")
    if verbose:
        print("Raw Text Sequence: {}".format(text))
        print("Phoneme Sequence: {}".format(s))
    return np.array(text_to_sequence(s, preprocess_config["preprocessing"]["text"]["text_cleaners"]))
