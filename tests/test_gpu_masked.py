"""Masked-documents fusion: the device bitmap AND-NOT at emission vs the reference's
masked_documents_registry::test() in the exec Handlers (exec.cpp:1108-1116).  Registry built by the reference's own pack/unpack."""
import numpy as np
import pytest

import trinity_b200 as tb
from util import Pair, assert_close_scores, assert_same_docs, assert_topk_equal, closed_form_lists

pytestmark = pytest.mark.gpu
NDOCS = 300_000
QUERIES = ["t1 AND t2", "t3 OR t7 OR t9", "t1 AND (t2 OR t3) NOT t5", "t10", "(t1 AND t2) OR (t3 AND t4)"]


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE], ids=["google", "lucene"])
def test_masked_documents_match_reference(ref, codec):
    p = Pair(ref, codec, closed_form_lists(NDOCS), NDOCS)
    rng = np.random.default_rng(3)
    masked = np.unique(np.concatenate([rng.integers(1, NDOCS + 1, 20_000), np.arange(6, 6000, 6), [1, NDOCS, 16384, 16383, 8192]])).astype(np.uint32)
    p.gpu.set_masked_documents(masked)
    res = p.gpu.exec_batch([p.plan(q) for q in QUERIES], tb.MODE_DOCS_ONLY)
    sres = p.gpu.exec_batch([p.plan(q, scored=True) for q in QUERIES], tb.MODE_SCORED_ALL)
    tres = p.gpu.exec_batch([p.plan(q, scored=True) for q in QUERIES], tb.MODE_SCORED_TOPK, k=50)
    for i, q in enumerate(QUERIES):
        want, _ = p.ref.exec_masked(q, False, masked, NDOCS + 1)
        unmasked, _ = p.ref.exec(q, False, NDOCS + 1)
        assert len(want) < len(unmasked)
        assert_same_docs(res.query(i)[0], want, f"[{q}] masked")
        assert int(res.match_counts[i]) == len(want)
        wd, ws = p.ref.exec_masked(q, True, masked, NDOCS + 1)
        gd, gs = sres.query(i)
        assert_same_docs(gd, wd, f"[{q}] masked scored")
        assert_close_scores(gs, ws, f"[{q}]")
        td, ts = tres.query(i)
        assert_topk_equal(td, ts, wd, ws, 50, f"[{q}] masked top-50")
    # clearing the registry restores the unmasked results
    p.gpu.set_masked_documents(None)
    res = p.gpu.exec_batch([p.plan(q) for q in QUERIES], tb.MODE_DOCS_ONLY)
    for i, q in enumerate(QUERIES):
        assert_same_docs(res.query(i)[0], p.ref.exec(q, False, NDOCS + 1)[0], f"[{q}] unmasked")
