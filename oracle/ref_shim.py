"""Import the UNMODIFIED reference ``layers`` package.

TEST / BASELINE INFRASTRUCTURE.  Two locations, in this order: ``oracle/_ref/`` (the staged copy made by
``oracle/build_ref.py``; git-ignored, travels to the GPU box like the built ``.so``) and ``/root/reference``
(build container only -- it does not exist on the GPU box).  Only ``tests/``, ``bench.py``'s CPU legs and the
scripts under ``oracle/`` import this module.  The two stubbed modules are text-front-end dependencies the
acoustic path never calls.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")
_MOUNTED = os.environ.get("ES_REFERENCE_DIR", "/root/reference")


def _pick() -> str:
    for d in (_STAGED, _MOUNTED):
        if os.path.isfile(os.path.join(d, "layers", "networks.py")):
            return d
    return _MOUNTED


REF_DIR = _pick()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "layers", "networks.py"))


def import_reference_layers():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_DIR}")
    for n in ("unidecode", "inflect"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["inflect"].engine = lambda: types.SimpleNamespace(number_to_words=lambda *a, **k: "")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import layers  # noqa: E402  (the reference package)
    return layers


def build_reference_model(cfg, state):
    """Reference Phoneme2Mel with our weights loaded strict=True (proves state-dict compatibility)."""
    import torch
    layers = import_reference_layers()
    enc = layers.PhonemeEncoder(pitch_stats=cfg.pitch_stats, energy_stats=cfg.energy_stats,
                                depth=cfg.depth, reduction=cfg.reduction, head=cfg.head,
                                embed_dim=cfg.embed_dim, kernel_size=cfg.kernel_size,
                                expansion=cfg.expansion)
    dec = layers.MelDecoder(dim=cfg.embed_dim // cfg.reduction, kernel_size=cfg.decoder_kernel_size,
                            n_blocks=cfg.n_blocks, block_depth=cfg.block_depth)
    m = layers.Phoneme2Mel(enc, dec).eval()
    m.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in state.items()}, strict=True)
    return m


def run_reference(model, batch, train):
    """batch: dict of numpy arrays in the reference layout -> dict of numpy outputs."""
    import torch
    tb = {k: torch.from_numpy(v.copy()) for k, v in batch.items()}
    with torch.no_grad():
        out = model(tb, train=train)
    if not train:
        mel, mel_len, dur = out
        return {"mel": mel.numpy(), "mel_len": mel_len.numpy(), "duration": dur.numpy()}
    return {k: (v.numpy() if v is not None else None) for k, v in out.items()}


def reference_functions(rel_path, names):
    """The UNMODIFIED source of the named top-level functions / methods of a reference file, compiled into a fresh
    namespace (numpy + torch only).  For files whose import needs packages this image lacks (datamodule.py imports
    lightning, utils/tools.py matplotlib): their pure functions still run as written."""
    import ast
    import numpy as np
    import torch
    root = _MOUNTED if os.path.isfile(os.path.join(_MOUNTED, rel_path)) else REF_DIR
    path = os.path.join(root, rel_path)
    if not os.path.isfile(path):
        raise RuntimeError(f"{rel_path} not found under the reference tree")
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            code = ast.get_source_segment(src, node)
            import textwrap
            exec(compile(textwrap.dedent(code), f"{rel_path}:{node.name}", "exec"), ns)
            found[node.name] = ns[node.name]
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"{rel_path}: no function(s) {sorted(missing)}")
    return found, ns
