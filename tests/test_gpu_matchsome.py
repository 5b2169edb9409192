"""MatchSome / min-should-match: [a, b, ...] with match_some.min = m -> DocsSetIterators::DisjunctionSome
(docset_iterators.cpp:679-811), the span of build_span (exec.cpp:453-466) — GPU bit-sliced counters vs the reference exec_query."""
import numpy as np
import pytest

import trinity_b200 as tb
from util import Pair, assert_close_scores, assert_same_docs, assert_topk_equal, closed_form_lists

pytestmark = pytest.mark.gpu
NDOCS = 300_000
QUERIES = [("[t1, t2, t3]", 2), ("[t1, t2, t3, t4, t5]", 3), ("[t1, t2, t3]", 1), ("[t1, t2, t3]", 3), ("[t1, t2]", 3),
           ("[t3, t4 AND t5, t6 OR t7, t2]", 2), ("t1 AND [t2, t3, t4]", 2), ("[t2, t3, t4] NOT t1", 2), ("[t2, t3, t4] OR t9", 2),
           ("[t2, nosuchterm, t4, t5]", 2), ("[t1, t2, t3, t4, t5, t6, t7, t8, t9, t10]", 5), ("[t1, t2, t3, t4, t5, t6, t7, t8, t9, t10]", 9),
           ("(t1 AND t2) OR [t3, t4, t5]", 2)]


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE], ids=["google", "lucene"])
def test_match_some_matches_reference(ref, codec):
    p = Pair(ref, codec, closed_form_lists(NDOCS), NDOCS)
    plans = [tb.parse_query(q, p.tdict, min_match=m) for q, m in QUERIES]
    res = p.gpu.exec_batch(plans, tb.MODE_DOCS_ONLY)
    for i, (q, m) in enumerate(QUERIES):
        want, _ = p.ref.exec(q, False, NDOCS + 1, parser_flags=16, min_match=m)
        assert_same_docs(res.query(i)[0], want, f"[{q}] min={m}")
        assert int(res.match_counts[i]) == len(want)
    splans = [p.gpu.set_bm25_weights(tb.parse_query(q, p.tdict, min_match=m), NDOCS) for q, m in QUERIES]
    sres = p.gpu.exec_batch(splans, tb.MODE_SCORED_ALL)
    tres = p.gpu.exec_batch(splans, tb.MODE_SCORED_TOPK, k=40)
    for i, (q, m) in enumerate(QUERIES):
        wd, ws = p.ref.exec(q, True, NDOCS + 1, parser_flags=16, min_match=m)
        gd, gs = sres.query(i)
        assert_same_docs(gd, wd, f"[{q}] min={m} scored")
        assert_close_scores(gs, ws, f"[{q}] min={m}")
        td, ts = tres.query(i)
        assert_topk_equal(td, ts, wd, ws, 40, f"[{q}] min={m} top-40")
