"""Full-size (100M docs) parity diagnosis: GPU and2 batch vs the reference exec_query vs the true intersection of the raw postings."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import trinity_b200 as tb
from bench import gen_queries
from refharness import RefIndex, load_ref

N, V, NQ = 100_000_000, 4096, int(os.environ.get("NQ", "1000"))
synth = tb.SynthIndex(tb.CODEC_GOOGLE, N, V, threads=os.cpu_count())
texts, ranks = gen_queries("and2", NQ, V)
g = tb.GpuIndexSource(0)
g.upload(synth.codec, np.asarray(synth.index), np.asarray(synth.terms), N)
tdict = tb.TermDictionary(synth.names)
plans = [tb.parse_query(t, tdict) for t in texts]
res = g.exec_batch(plans, tb.MODE_DOCS_ONLY)
r = RefIndex.from_bytes(load_ref(), synth.codec, np.asarray(synth.index), np.asarray(synth.hits), synth.names, np.asarray(synth.terms), N, synth.sum_hits)
el, counts, sums, _, _ = r.exec_batch(texts, False, 100, os.cpu_count())
bad = np.flatnonzero(np.asarray(res.match_counts, np.uint64) != counts)
print("mismatching queries:", len(bad), "of", NQ)
for q in bad[:12]:
    a, b = int(ranks[q][0]) + 1, int(ranks[q][1]) + 1
    da, _ = tb.SynthIndex.postings(N, a)
    db, _ = tb.SynthIndex.postings(N, b)
    true = np.intersect1d(da, db, assume_unique=True)
    gd = res.query(int(q))[0]
    rd, _ = r.exec(texts[q], False, max(len(true), int(counts[q])) + 10)
    print(f"q{q} ranks ({a},{b}) df ({len(da)},{len(db)}): gpu {int(res.match_counts[q])} ref {int(counts[q])} true {len(true)}  gpu==true {np.array_equal(gd, true)} ref==true {np.array_equal(rd, true)}")
    if not np.array_equal(rd, true):
        miss = np.setdiff1d(true, rd); extra = np.setdiff1d(rd, true)
        print("   ref misses", len(miss), "first", miss[:5], "extra", len(extra), extra[:5])
    if not np.array_equal(gd, true):
        miss = np.setdiff1d(true, gd); extra = np.setdiff1d(gd, true)
        print("   gpu misses", len(miss), "first", miss[:5], "extra", len(extra), extra[:5])
