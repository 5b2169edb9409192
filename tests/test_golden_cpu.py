"""Golden vectors (generated from the reference, tests/golden/make_golden.py) vs: our host encoders, the load-time directory and
the plain-C oracle.  Runs without /root/reference and without oracle/_ref."""
import numpy as np
import pytest

import oracle_c
import trinity_b200 as tb
from golden_util import check_docs_digest, check_scores_digest, load_closed, load_lists
from util import closed_form_lists

CODECS = [tb.CODEC_GOOGLE, tb.CODEC_LUCENE]


@pytest.mark.parametrize("codec", CODECS)
def test_our_encoders_reproduce_golden_bytes(codec):
    z, n = load_lists(codec)
    b = tb.IndexBuilder(codec)
    for i in range(n):
        b.add_term(z[f"docids_{i}"], z[f"infreqs_{i}"], z[f"positions_{i}"])
    assert np.array_equal(b.terms_array(), z["terms"])
    mine, gold = b.index(), z["index"]
    diff = np.flatnonzero(mine != gold)
    # GOOGLE: identical. LUCENE: identical except the reference's uninitialised FastPFor padding bytes (ours are 0)
    assert np.all(mine[diff] == 0) and (codec == tb.CODEC_LUCENE or diff.size == 0) and diff.size < mine.size // 50
    if codec == tb.CODEC_LUCENE:
        hm, hg = b.hits(), z["hits"]
        hd = np.flatnonzero(hm != hg)
        assert hm.size == hg.size and np.all(hm[hd] == 0)


@pytest.mark.parametrize("codec", CODECS)
def test_c_oracle_decodes_golden_bytes(codec):
    z, n = load_lists(codec)
    L = oracle_c.load()
    for i in range(n):
        d, f = oracle_c.decode(L, codec, z["index"], z["terms"][i])
        assert np.array_equal(d, z[f"docids_{i}"]) and np.array_equal(f, z[f"freqs_{i}"])


@pytest.mark.parametrize("codec", CODECS)
def test_directory_on_golden_bytes(codec):
    z, n = load_lists(codec)
    bs = 32 if codec == tb.CODEC_GOOGLE else 128
    for i in range(n):
        d = z[f"docids_{i}"]
        last, off, first = tb.directory_probe(codec, z["index"], tuple(z["terms"][i]))
        assert first == d[0] and np.array_equal(last[:-1], d[np.minimum(np.arange(bs, len(d) + bs, bs), len(d)) - 1])


@pytest.mark.parametrize("codec", CODECS)
def test_c_oracle_exec_reproduces_golden_results(codec):
    z = load_closed(codec)
    ndocs = int(z["ndocs"][0])
    lists = closed_form_lists(ndocs)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f)
    names = [f"t{i + 1}" for i in range(len(lists))]
    tdict = tb.TermDictionary(names)
    L = oracle_c.load()
    for qi, q in enumerate(z["queries"].tolist()):
        nodes = tb.parse_query(q, tdict)
        for x in nodes:
            if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                x["weight"] = tb.bm25_idf(len(lists[int(x["term"])][0]), ndocs)
        ids, _ = oracle_c.exec_query(L, codec, b.index(), b.terms_array(), nodes, ndocs, False)
        check_docs_digest(z, qi, ids, f"[{q}]")
        if f"ssum_{qi}" in z:
            sid, sc = oracle_c.exec_query(L, codec, b.index(), b.terms_array(), nodes, ndocs, True)
            check_docs_digest(z, qi, sid, f"[{q}] scored")
            check_scores_digest(z, qi, sid, sc, f"[{q}]")


@pytest.mark.parametrize("codec", CODECS)
def test_c_oracle_reproduces_widened_golden_results(codec):
    """rows added after SURVEY 8a-e, against fixtures generated from the reference: MatchSome groups on the closed-form index and phrase
    plans on the reference-encoded document-major corpus (index + hits.data bytes are part of the fixture)"""
    from golden_util import load_widened
    z = load_widened(codec)
    L = oracle_c.load()
    ndocs = int(z["ndocs"][0])
    lists = closed_form_lists(ndocs)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f)
    tdict = tb.TermDictionary([f"t{i + 1}" for i in range(len(lists))])
    for qi, (q, m) in enumerate(zip(z["some_queries"].tolist(), z["some_min"].tolist())):
        nodes = tb.parse_query(q, tdict, min_match=int(m))
        for x in nodes:
            if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                x["weight"] = tb.bm25_idf(len(lists[int(x["term"])][0]), ndocs)
        sid, sc = oracle_c.exec_query(L, codec, b.index(), b.terms_array(), nodes, ndocs, True)
        check_docs_digest(z, f"s{qi}", sid, f"[{q}] min={m}")
        check_scores_digest(z, f"s{qi}", sid, sc, f"[{q}] min={m}")
    pn = int(z["phrase_ndocs"][0])
    terms = z["phrase_terms"]
    pdict = tb.TermDictionary([f"w{t + 1}" for t in range(len(terms))])
    for qi, q in enumerate(z["phrase_queries"].tolist()):
        nodes = tb.parse_query(q, pdict)
        for x in nodes:
            if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
                x["weight"] = tb.bm25_idf(int(terms["documents"][x["term"]]), pn)
        sid, sc = oracle_c.exec_query(L, codec, z["phrase_index"], terms, nodes, pn, True, hits=z["phrase_hits"])
        check_docs_digest(z, f"p{qi}", sid, f"[{q}]")
        check_scores_digest(z, f"p{qi}", sid, sc, f"[{q}]")
