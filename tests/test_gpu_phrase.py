"""Positions on the device: phrase plans through the C ABI vs the reference's exec_query — GOOGLE codec (inline hits,
google_codec.cpp:533-594 materialize_hits + Phrase::consider_phrase_match docset_iterators.cpp:66-158).  The corpus is generated
DOCUMENT-major (one term per position: the reference's DocWordsSpace keeps one term per position).  Documents bit-exact in DocumentsOnly and
scored mode, phrase scores (score(matchCnt, sum idf), docset_iterators_scorers.cpp:195-228) within 1e-5, top-k per assert_topk_equal.
The LUCENE codec keeps its hits in hits.data (lucene_codec.cpp:401-513, :767-856): executed once trn_upload_hits has handed it over; without
it phrase plans are refused, loudly."""
import numpy as np
import pytest

import trinity_b200 as tb
from refharness import RefIndex
from test_phrase_cpu import QUERIES
from util import assert_close_scores, assert_same_docs, assert_topk_equal

pytestmark = pytest.mark.gpu


def _corpus(ref, codec, ndocs, vocab, seed, minlen, maxlen, with_hits=True):
    rng = np.random.default_rng(seed)
    prob = 1.0 / np.arange(1, vocab + 1)
    prob /= prob.sum()
    per_term = [dict() for _ in range(vocab)]
    for d in range(1, ndocs + 1):
        toks = rng.choice(vocab, size=int(rng.integers(minlen, maxlen)), p=prob)
        for pos, t in enumerate(toks, start=1):
            per_term[int(t)].setdefault(d, []).append(pos)
    names = [f"w{t + 1}" for t in range(vocab)]
    r = RefIndex(ref, codec)
    b = tb.IndexBuilder(codec)
    for t in range(vocab):
        docs = np.array(sorted(per_term[t]), np.uint32)
        freqs = np.array([len(per_term[t][int(d)]) for d in docs], np.uint32)
        flat = np.array([p for d in docs for p in per_term[t][int(d)]], np.uint32)
        r.add_term(names[t], docs, freqs, flat)
        b.add_term(docs, freqs, flat)
    r.finish(ndocs)
    g = tb.GpuIndexSource(0)
    g.upload(codec, b.index(), b.terms_array(), ndocs)
    if codec == tb.CODEC_LUCENE and with_hits:
        g.upload_hits(b.index(), b.hits())
    return r, g, tb.TermDictionary(names)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE], ids=["google", "lucene"])
@pytest.mark.parametrize("shape", ["short", "long"])
def test_phrases_match_reference(ref, shape, codec):
    # "long": documents of up to 300 tokens over 9 terms => hundreds of hits per (term, document): the 64-position chunks of the checker,
    # and (LUCENE) runs of hits that cross 128-hit blocks and reach into the varbyte tail
    ndocs, vocab, lo, hi = (60_000, 9, 3, 30) if shape == "short" else (3_000, 9, 80, 300)
    r, g, tdict = _corpus(ref, codec, ndocs, vocab, 21, lo, hi)
    qs = [q for q in QUERIES]
    plans = [tb.parse_query(q, tdict) for q in qs]
    res = g.exec_batch(plans, tb.MODE_DOCS_ONLY)
    splans = [g.set_bm25_weights(tb.parse_query(q, tdict), ndocs) for q in qs]
    sres = g.exec_batch(splans, tb.MODE_SCORED_ALL)
    tres = g.exec_batch(splans, tb.MODE_SCORED_TOPK, k=25)
    nonempty = 0
    for i, q in enumerate(qs):
        want, _ = r.exec(q, False, ndocs + 1)
        assert_same_docs(res.query(i)[0], want, f"[{q}] {shape}")
        assert int(res.match_counts[i]) == len(want)
        nonempty += len(want) > 0
        wd, ws = r.exec(q, True, ndocs + 1)
        gd, gs = sres.query(i)
        assert_same_docs(gd, wd, f"[{q}] {shape} scored")
        assert_close_scores(gs, ws, f"[{q}] {shape}")
        td, ts = tres.query(i)
        assert_topk_equal(td, ts, wd, ws, 25, f"[{q}] {shape} top-25")
    assert nonempty >= 9
    g.close()


def test_phrases_are_refused_on_lucene_without_its_hits(ref):
    r, g, tdict = _corpus(ref, tb.CODEC_LUCENE, 2_000, 9, 21, 3, 20, with_hits=False)
    with pytest.raises(tb.TrinityError, match="rc=-7"):
        g.exec_batch([tb.parse_query('"w1 w2"', tdict)], tb.MODE_DOCS_ONLY)
    res = g.exec_batch([tb.parse_query("w1 AND w2", tdict)], tb.MODE_DOCS_ONLY)  # everything else keeps working
    assert_same_docs(res.query(0)[0], r.exec("w1 AND w2", False, 2001)[0], "lucene and")
    g.close()
