"""Full-size checksum diagnosis: copy=True vs copy=False result views, offsets vs counts, per-query docID checksums vs the reference."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trinity_b200 as tb
from bench import gen_queries
from refharness import RefIndex, load_ref

N, V, NQ = 100_000_000, 4096, 1000
synth = tb.SynthIndex(tb.CODEC_GOOGLE, N, V, threads=os.cpu_count())
texts, ranks = gen_queries("and2", NQ, V)
g = tb.GpuIndexSource(0)
g.upload(synth.codec, np.asarray(synth.index), np.asarray(synth.terms), N)
tdict = tb.TermDictionary(synth.names)
plans = [tb.parse_query(t, tdict) for t in texts]
r = RefIndex.from_bytes(load_ref(), synth.codec, np.asarray(synth.index), np.asarray(synth.hits), synth.names, np.asarray(synth.terms), N, synth.sum_hits)
el, counts, sums, _, _ = r.exec_batch(texts, False, 100, os.cpu_count())

def check(res, tag):
    off = np.asarray(res.offsets, np.int64)
    print(tag, "offsets monotone", bool(np.all(np.diff(off) >= 0)), "diff==counts", bool(np.array_equal(np.diff(off).astype(np.uint64), np.asarray(res.match_counts, np.uint64))),
          "counts==ref", bool(np.array_equal(np.asarray(res.match_counts, np.uint64), counts)), "total", int(off[-1]))
    bad = []
    for q in range(NQ):
        d = res.docids[off[q]:off[q + 1]]
        if int(d.astype(np.uint64).sum()) != int(sums[q]):
            bad.append(q)
    print(tag, "checksum mismatches:", len(bad), bad[:10])
    for q in bad[:3]:
        rd, _ = r.exec(texts[q], False, int(counts[q]) + 10)
        d = res.docids[off[q]:off[q + 1]]
        print("   q", q, texts[q], "n", len(d), "sorted", bool(np.all(np.diff(d.astype(np.int64)) > 0)), "equal to ref", bool(np.array_equal(d, rd)),
              "first diff", int(np.flatnonzero(d != rd)[0]) if not np.array_equal(d, rd) and len(d) == len(rd) else None)
    # whole-array checksum the way bench.py does it
    cs = np.concatenate([[0], np.cumsum(np.asarray(res.docids[: off[-1]], np.uint64), dtype=np.uint64)])
    got = cs[off[1:]] - cs[off[:-1]]
    print(tag, "cumsum-based equal:", bool(np.array_equal(got, sums)), "mismatches", int(np.count_nonzero(got != sums)))

check(g.exec_batch(plans, tb.MODE_DOCS_ONLY), "copy=True ")
packed = g.pack(plans)
res = g.exec_batch(plans, tb.MODE_DOCS_ONLY, 100, copy=False, packed=packed)
check(res, "copy=False")
