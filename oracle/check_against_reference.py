"""Pin the numpy oracle to the unmodified reference (run in the build container).

    python oracle/check_against_reference.py

For tiny/small/base x {teacher-forced, free-running} x {B>1 ragged, B==1} this runs
``layers.Phoneme2Mel`` (imported from /root/reference, weights loaded strict=True) and
``oracle.es_oracle.phoneme2mel`` on the same seeded weights and inputs and prints the
max-abs differences.  Integer outputs (mel_len, masks) must be identical.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientspeech_b200.config import VARIANTS  # noqa: E402
from efficientspeech_b200.params import init_state_dict  # noqa: E402
from efficientspeech_b200.synthetic import make_batch  # noqa: E402
from oracle import es_oracle  # noqa: E402
from oracle.ref_shim import build_reference_model, run_reference  # noqa: E402


def main():
    worst = 0.0
    for name, cfg in VARIANTS.items():
        sd = init_state_dict(cfg, seed=1)
        ref = build_reference_model(cfg, sd)
        for (B, N, ragged) in [(4, 37, True), (3, 64, True), (1, 50, False), (2, 1, False), (5, 6, True)]:
            batch = make_batch(cfg, B, N, seed=B * 100 + N, ragged=ragged, fixed_duration=None)
            for train in (True, False):
                try:
                    r = run_reference(ref, batch, train)
                except Exception as e:  # e.g. T == 0 raises inside the reference conv
                    print(f"{name} B={B} N={N} train={train}: reference raised {type(e).__name__}: {e}")
                    continue
                o = es_oracle.phoneme2mel(batch, sd, train=train)
                assert np.array_equal(r["mel_len"], o["mel_len"]), (r["mel_len"], o["mel_len"])
                assert r["mel"].shape == o["mel"].shape, (r["mel"].shape, o["mel"].shape)
                dm = float(np.abs(r["mel"] - o["mel"]).max())
                dd = float(np.abs(r["duration"] - o["duration"]).max())
                extra = ""
                if train:
                    dp = float(np.abs(r["pitch"] - o["pitch"]).max())
                    de = float(np.abs(r["energy"] - o["energy"]).max())
                    df = float(np.abs(r["features"] - o["features"]).max())
                    if r["masks"] is None:
                        assert o["masks"] is None
                    else:
                        assert np.array_equal(r["masks"], o["masks"])
                    extra = f" pitch {dp:.2e} energy {de:.2e} features {df:.2e}"
                    worst = max(worst, dp, de, df)
                worst = max(worst, dm, dd)
                print(f"{name:5s} B={B} N={N:3d} train={int(train)} T={o['mel'].shape[1]:4d} "
                      f"mel {dm:.2e} dur {dd:.2e}{extra}")
    print(f"worst max-abs difference oracle vs reference: {worst:.3e}")
    assert worst < 5e-5, worst
    print("ORACLE PINNED TO REFERENCE: OK")


if __name__ == "__main__":
    main()
