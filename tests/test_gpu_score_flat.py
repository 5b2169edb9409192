"""k_score_flat (scored flat disjunctions on the LUCENE codec: DocsSetSpanForDisjunctionsWithThreshold + Scorer::score + the top-k sink,
docset_spans.cpp:681-790, similarity.h:228-235) against the reference's exec_query on postings with freq 0 (Scorer::score(0) == 0: the
document matches with score +0, which the -0.0f "untouched" sentinel of the score tile must tell apart), freq >= 64 (outside the per-term
table), tail blocks, single terms, documents matched by every term and by one, sparse-only disjunctions (most tiles see no posting), and
weights twelve orders of magnitude apart (a term that every document holds has an idf ~1e-6)."""
import numpy as np
import pytest

import trinity_b200 as tb
from util import Pair, assert_close_scores, assert_same_docs, assert_topk_equal

pytestmark = pytest.mark.gpu
NDOCS = 150_000


def _lists():
    rng = np.random.default_rng(12)
    out = []
    # dense, medium, sparse terms; freqs 0 (one posting in 7), 1..5, and a few >= 64
    for n in (70_000, 40_000, 9_000, 3_001, 640, 129, 127, 50, 20_000, 11_111, 5_000, 2_500):
        d = np.sort(rng.choice(np.arange(1, NDOCS + 1, dtype=np.uint32), size=n, replace=False))
        f = rng.integers(1, 6, n).astype(np.uint32)
        f[rng.random(n) < 1 / 7] = 0
        f[rng.random(n) < 0.01] = rng.integers(64, 3000)
        out.append((d, f))
    out.append((np.arange(1, NDOCS + 1, dtype=np.uint32), np.ones(NDOCS, np.uint32)))  # t13: every document (idf ~ 3e-6)
    return out


QUERIES = [
    "t1",
    "t8",
    "t1 OR t2",
    "t3 OR t5 OR t7",
    " OR ".join(f"t{i}" for i in range(1, 11)),
    " OR ".join(f"t{i}" for i in range(1, 13)),
    "t6 OR t7 OR t8",          # sparse terms only: most tiles see no posting
]
EVERY_DOC = ["t13 OR t3", "t13"]  # the every-document term: weights ~1e-6 next to ~5


def test_flat_scored_disjunctions_match_reference(ref):
    p = Pair(ref, tb.CODEC_LUCENE, _lists(), NDOCS)
    saw_zero = False
    for batch in (QUERIES, QUERIES + EVERY_DOC, EVERY_DOC):
        plans = [p.plan(q, scored=True) for q in batch]
        allres = p.gpu.exec_batch(plans, tb.MODE_SCORED_ALL)
        top = {k: p.gpu.exec_batch(plans, tb.MODE_SCORED_TOPK, k=k) for k in (10, 100)}
        for i, q in enumerate(batch):
            wd, ws = p.ref.exec(q, True, NDOCS + 1)
            saw_zero |= bool((ws == 0).any())
            gd, gs = allres.query(i)
            assert_same_docs(gd, wd, f"[{q}] scored-all")
            assert_close_scores(gs, ws, f"[{q}] scored-all")
            for k, res in top.items():
                td, ts = res.query(i)
                assert int(res.match_counts[i]) == len(wd), f"[{q}] match count"
                assert_topk_equal(td, ts, wd, ws, k, f"[{q}] top-{k}")
    assert saw_zero, "the corpus is meant to hold documents that match with score 0 (freq-0 postings)"
    p.gpu.close()
