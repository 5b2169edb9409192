#!/usr/bin/env python
"""configs[4]: postings-decode microbench — whole-list decode of every term of the synthetic index (k_decode_terms),
materialised (8 B/posting written) and fused (checksum only), both codecs, vs the HBM roofline."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import trinity_b200 as tb

ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
for codec, name in ((0, "GOOGLE"), (1, "LUCENE")):
    s = tb.SynthIndex(codec, ndocs, 4096)
    g = tb.GpuIndexSource(0)
    g.upload(codec, np.asarray(s.index), np.asarray(s.terms), ndocs)
    terms = list(range(4096))
    postings = int(s.terms["documents"].sum())
    chunk_bytes = int(s.terms["chunk_len"].sum())
    for mat in (False, True):
        if mat and postings * 8 > 20e9:
            continue
        best = 1e9
        for it in range(5):
            if mat:
                # materialised variant: time only the kernel (device_ms); D2H of 8 B/posting is outside the event pair
                d, f, sums, ms = g.decode_terms(terms[:512] if postings > 5e8 else terms, materialise=True)
            else:
                d, f, sums, ms = g.decode_terms(terms, materialise=False)
            if it >= 2:
                best = min(best, ms)
        p = postings if not mat else int(s.terms["documents"][:512].sum()) if postings > 5e8 else postings
        cb = chunk_bytes if not mat or postings <= 5e8 else int(s.terms["chunk_len"][:512].sum())
        algo = cb + (8 * p if mat else 0)
        print(json.dumps({"codec": name, "variant": "materialised" if mat else "fused-checksum", "ndocs": ndocs, "postings": p,
                          "kernel_ms": best, "postings_per_s": p / (best * 1e-3), "algorithmic_bytes": algo,
                          "achieved_gbs": algo / (best * 1e-3) / 1e9, "peak_gbs": peak, "frac": algo / (best * 1e-3) / 1e9 / peak,
                          "bytes_per_posting": cb / p}))
    g.close()
