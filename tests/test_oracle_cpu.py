"""Pins the plain-C restatement (oracle/trinity_oracle.c) against the reference's own code (oracle/_ref): decode streams,
BM25, and exec_query result sets + scores.  CPU only."""
import numpy as np
import pytest

import oracle_c
import trinity_b200 as tb
from refharness import RefIndex
from test_codecs_cpu import make_lists, positions_for
from test_frontend_cpu import EXTRA
from test_gpu_parity import TEMPLATES
from util import closed_form_lists

NDOCS = 30_000


@pytest.fixture(scope="module")
def orc():
    return oracle_c.load()


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_restated_decoders_match_reference_iterators(ref, orc, codec):
    rng = np.random.default_rng(11 + codec)
    lists = make_lists(rng)
    r = RefIndex(ref, codec)
    for i, (d, f) in enumerate(lists):
        r.add_term(f"t{i}", d, f, positions_for(f, rng))
    r.finish(int(max(int(d[-1]) for d, _ in lists)))
    index, terms = r.index(), r.terms()
    for i, (d, f) in enumerate(lists):
        rd, rf = r.decode(i, len(d) + 4)
        od, of = oracle_c.decode(orc, codec, index, terms[i])
        assert np.array_equal(od, rd) and np.array_equal(of, rf)


def test_restated_bm25_matches_reference(ref, orc):
    r = RefIndex(ref, tb.CODEC_GOOGLE)
    lists = closed_form_lists(NDOCS)
    for i, (d, f) in enumerate(lists):
        r.add_term(f"t{i + 1}", d, f)
    r.finish(NDOCS)
    for t in range(len(lists)):
        idf = orc.orc_bm25_idf(len(lists[t][0]), NDOCS)
        assert abs(idf - tb.bm25_idf(len(lists[t][0]), NDOCS)) < 1e-12
        for fr in (0, 1, 3, 8, 100, 65535):
            a, b = orc.orc_bm25_score(idf, fr), r.bm25(t, fr)
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-30)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_restated_exec_matches_reference_exec_query(ref, orc, codec):
    lists = closed_form_lists(NDOCS)
    names = [f"t{i + 1}" for i in range(len(lists))]
    r = RefIndex(ref, codec)
    for n, (d, f) in zip(names, lists):
        r.add_term(n, d, f)
    r.finish(NDOCS)
    index, terms = r.index(), r.terms()
    tdict = tb.TermDictionary(names)
    for q in TEMPLATES + EXTRA:
        nodes = tb.parse_query(q, tdict)
        for x in nodes:
            if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                x["weight"] = tb.bm25_idf(int(terms["documents"][x["term"]]), NDOCS)
        want, _ = r.exec(q, False, NDOCS + 1)
        got, _ = oracle_c.exec_query(orc, codec, index, terms, nodes, NDOCS, False)
        assert np.array_equal(got, want), q
        if "nosuchterm" in q:
            continue
        wd, ws = r.exec(q, True, NDOCS + 1)
        gd, gs = oracle_c.exec_query(orc, codec, index, terms, nodes, NDOCS, True)
        assert np.array_equal(gd, wd), q
        rel = np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-30)
        assert rel.max() <= 1e-5, q
    # MatchSome groups (DisjunctionSome)
    for q, m in (("[t1, t2, t3]", 2), ("[t3, t4 AND t5, t6 OR t7, t2]", 2), ("t1 AND [t2, t3, t4, t5]", 3), ("[t2, t3, t4] OR t9", 2)):
        nodes = tb.parse_query(q, tdict, min_match=m)
        for x in nodes:
            if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                x["weight"] = tb.bm25_idf(int(terms["documents"][x["term"]]), NDOCS)
        wd, ws = r.exec(q, True, NDOCS + 1, parser_flags=16, min_match=m)
        gd, gs = oracle_c.exec_query(orc, codec, index, terms, nodes, NDOCS, True)
        assert np.array_equal(gd, wd), q
        rel = np.abs(gs - ws) / np.maximum(np.abs(ws), 1e-30)
        assert rel.max() <= 1e-5, q


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_restated_positions_match_reference_materialize_hits(ref, orc, codec):
    """hit streams (Google inline hits, Lucene hits.data incl. 128-hit PFor blocks and the varbyte tail) restated in C == the positions the
    reference's PostingsListIterator::materialize_hits yields, == the positions that went in"""
    rng = np.random.default_rng(31 + codec)
    r = RefIndex(ref, codec)
    want = []
    for t, (ndocs_t, maxf) in enumerate(((40, 3), (700, 9), (3000, 2), (129, 40), (1, 1), (260, 1))):
        docs = np.sort(rng.choice(np.arange(1, 50_000), size=ndocs_t, replace=False)).astype(np.uint32)
        freqs = rng.integers(1, maxf + 1, size=ndocs_t).astype(np.uint32)
        pos = np.concatenate([np.cumsum(rng.integers(1, 40, size=int(f))) for f in freqs]).astype(np.uint32)
        r.add_term(f"p{t}", docs, freqs, pos)
        want.append(pos)
    r.finish(50_000)
    index, hits, terms = r.index(), r.hits(), r.terms()
    for t, pos in enumerate(want):
        got_ref = r.positions(t, len(pos) + 8)
        assert np.array_equal(got_ref, pos), f"reference positions of term {t}"
        got = oracle_c.positions(orc, codec, index, hits, terms[t])
        assert np.array_equal(got, pos), f"restated positions of term {t}"
