"""GPU-side GOOGLE encoder (trn_encode_google) vs the host encoder on the same postings: device time of the encode (kernels + scans, without
the host<->device copies), bytes identical.  Usage: python scripts/microbench_encode.py [ndocs] [first_rank] [nterms] [with_positions]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import trinity_b200 as tb  # noqa: E402


def main():
    ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    nterms = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    with_pos = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) != 0
    lists = []
    for rank in range(first, first + nterms):
        d, f = tb.SynthIndex.postings(ndocs, rank, 1000, 0x5EED)
        p = tb.SynthIndex.positions(ndocs, rank, 1000, 0x5EED) if with_pos else None
        lists.append((d, f, p))
    posts = sum(len(l[0]) for l in lists)
    hits = sum(int(l[1].sum()) for l in lists)
    t0 = time.perf_counter()
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    for d, f, p in lists:
        b.add_term(d, f, p)
    host_s = time.perf_counter() - t0
    want = b.index()
    g = tb.GpuIndexSource(0)
    best = None
    for _ in range(4):  # first call: allocations
        t0 = time.perf_counter()
        index, terms, _, ms = g.encode_google(lists)
        wall = time.perf_counter() - t0
        best = ms if best is None else min(best, ms)
    same = bool(index.size == want.size and np.array_equal(index, want) and np.array_equal(terms, b.terms_array()))
    in_bytes = posts * 8 + (hits * 4 if with_pos else 0) + (posts * 8 if with_pos else 0)  # docids + freqs (+ positions + the hit offsets)
    print(json.dumps({"what": "GOOGLE encode, device vs host", "ndocs": ndocs, "terms": [first, first + nterms - 1], "postings": posts, "hits": hits,
                      "with_positions": with_pos, "index_bytes": int(index.size), "bytes_identical_to_host_encoder": same,
                      "device_ms": round(best, 3), "postings_per_s_device": posts / (best / 1e3), "in_plus_out_GBps": (in_bytes + index.size) / (best / 1e3) / 1e9,
                      "call_wall_s_incl_copies": round(wall, 3), "host_encoder_s_1_thread": round(host_s, 3),
                      "host_postings_per_s": posts / host_s}))
    assert same


if __name__ == "__main__":
    main()
