"""Recipe: stage the UNMODIFIED reference modules of the acoustic path under ``oracle/_ref/``.

    python oracle/build_ref.py            (build container only: needs /root/reference)

TEST / BASELINE INFRASTRUCTURE.  ``oracle/_ref/`` is listed in ``.gitignore`` (reference sources never
enter this repository's history) but NOT in ``.gpurunignore``, so the staged copy travels to the GPU box
next to the built ``.so`` files.  There ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the
reference's own ``layers.Phoneme2Mel.forward`` (layers/networks.py:415) under torch on the host cores
(``kind: "reference"``), and the parity tests may use it as a second checker beside the numpy restatement.
The product package never imports it.

Files staged (Apache-2.0, LICENSE copied alongside):
  layers/{__init__,blocks,networks,acoustic}.py, text/*.py (imported by layers/networks.py:13),
  preprocessed_data/LJSpeech/stats.json (model.py:127-130), hifigan/{__init__,models}.py +
  hifigan/LJ_V2/{config.json,generator_v2} (vocoder row, model.py:23-48).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("ES_REFERENCE_DIR", "/root/reference")

FILES = [
    "LICENSE",
    "layers/__init__.py", "layers/blocks.py", "layers/networks.py", "layers/acoustic.py",
    "text/__init__.py", "text/cleaners.py", "text/cmudict.py", "text/numbers.py", "text/symbols.py",
    "text/tagdict.py",
    "preprocessed_data/LJSpeech/stats.json",
    "hifigan/__init__.py", "hifigan/models.py", "hifigan/LICENSE",
    "hifigan/LJ_V2/config.json", "hifigan/LJ_V2/generator_v2",
    # parsed, never imported (they need lightning / matplotlib): the sources of collate_fn, the loss and the schedule are
    # compiled function by function (oracle/ref_shim.reference_functions) to pin the rows next to the path
    "model.py", "datamodule.py", "utils/tools.py",
]


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "layers", "networks.py"))


def main() -> int:
    if not os.path.isfile(os.path.join(SRC, "layers", "networks.py")):
        print(f"reference tree not found at {SRC}; keeping whatever is staged ({'present' if staged() else 'absent'})")
        return 0
    n = 0
    for rel in FILES:
        s = os.path.join(SRC, rel)
        if not os.path.isfile(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        n += 1
    print(f"staged {n} reference files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
