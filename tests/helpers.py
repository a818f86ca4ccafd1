"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from efficientspeech_b200.config import VARIANTS
from efficientspeech_b200.params import init_state_dict, state_checksum

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances (max-abs) -- BASELINE.json north_star: mel within 1e-3 of the reference fp32 CPU
# path; SURVEY.md section 8c: scalar predictions within 1e-4; integers exact.
TOL_MEL = 1e-3
TOL_PRED = 1e-4


def golden_files():
    """Acoustic-path fixtures (<variant>_b<B>n<N>.npz); other rows keep their own fixtures (collate_*.npz, ...)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if os.path.basename(p).split("_")[0] in VARIANTS)


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    vname = str(g["variant"])
    cfg = VARIANTS[vname]
    sd = init_state_dict(cfg, seed=int(g["weight_seed"]))
    assert state_checksum(sd) == str(g["weight_checksum"]), "weights regenerated differently than at fixture time"
    batch = {k[3:]: g[k] for k in g if k.startswith("in_")}
    return vname, cfg, sd, batch, g
