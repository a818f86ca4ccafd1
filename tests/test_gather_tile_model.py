"""CPU model of the gathered x-tile staging of the first decoder layer (es_umma_dec.cu, GX kernels).

The issue warp stages the DISTINCT table rows a 64-frame tile (+2-frame halo) needs as one contiguous
copy (+ the padded-frame row) and writes a tile-row -> staged-row map; the producers read the conv window
through the map.  This restates that index arithmetic in numpy, lane by lane, and checks it against the
plain definition x[b, t] = P[rows[b, t]] (zero outside the utterance) for random durations, including
zero-duration runs that force the per-frame fallback, empty utterances and T not a multiple of 64."""
import numpy as np

TM, HALO, XROWS = 64, 2, 68


def frame_rows(dur, T):
    """es_gather.cu frame_source_kernel: rows[b,t] = b*N + upper_bound(cum[b], t) for t < mel_len[b], else B*N."""
    B, N = dur.shape
    cum = np.cumsum(dur, axis=1)
    rows = np.full((B, T), B * N, np.int64)
    for b in range(B):
        for t in range(min(T, int(cum[b, -1]))):
            rows[b, t] = b * N + int(np.searchsorted(cum[b], t, side="right"))
    return rows


def stage_tile(P, rows, b, t0, T, pad_id):
    """One tile as the issue warp stages it: returns (slot rows [68, C], map [68], compact?)."""
    lo, hi = max(t0 - HALO, 0), min(t0 + TM + HALO, T)
    head, nfr = lo - (t0 - HALO), hi - lo
    sx = np.full((32, 3), -1, np.int64)                      # lane l, k: frame lo + l + 32k
    for lane in range(32):
        for k in range(3):
            r = lo + lane + 32 * k
            if r < hi:
                sx[lane, k] = rows[b, r]
    first = sx[0, 0]
    mx = max([v for v in sx.reshape(-1) if v != pad_id] + [-1])
    has_pad = int((sx == pad_id).any())
    nr = mx - first + 1 if mx >= 0 else 0
    compact = nr + has_pad <= XROWS
    mp = np.full(XROWS, -7, np.int64)
    for r in range(XROWS):
        if r < head or r >= head + nfr:
            mp[r] = -1
    slot = np.full((XROWS, P.shape[1]), np.nan)
    for lane in range(32):
        for k in range(3):
            r = head + lane + 32 * k
            if sx[lane, k] >= 0:
                mp[r] = (nr if sx[lane, k] == pad_id else sx[lane, k] - first) if compact else r
                if not compact:
                    slot[r] = P[sx[lane, k]]
    if compact:
        if nr:
            slot[:nr] = P[first:first + nr]
        if has_pad:
            slot[nr] = P[pad_id]
    return slot, mp, compact


def test_gathered_tile_staging_matches_definition():
    rng = np.random.default_rng(5)
    seen = {True: 0, False: 0}
    for trial in range(60):
        B, N = int(rng.integers(1, 4)), int(rng.integers(1, 90)) + (120 if trial % 4 == 0 else 0)
        dur = rng.integers(0, 9, size=(B, N))
        if trial % 4 == 0:
            dur[:, 20:95] = 0                                # a 75-phoneme zero-duration run overflows the 68-row slot
            dur[:, 19] = dur[:, 95] = 3
        if trial % 9 == 0:
            dur[0] = 0
        T = int(dur.sum(1).max()) + int(rng.integers(0, 3))
        if T == 0:
            continue
        C = 4
        P = rng.standard_normal((B * N + 1, C))
        rows = frame_rows(dur, T)
        for b in range(B):
            for t0 in range(0, T, TM):
                slot, mp, compact = stage_tile(P, rows, b, t0, T, B * N)
                seen[compact] += 1
                assert (mp != -7).all()
                for r in range(XROWS):
                    t = t0 - HALO + r
                    want = P[rows[b, t]] if 0 <= t < T else np.zeros(C)
                    got = slot[mp[r]] if mp[r] >= 0 else np.zeros(C)
                    assert np.array_equal(got, want), (trial, b, t0, r)
    assert seen[True] > 0 and seen[False] > 0
