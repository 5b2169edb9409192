"""Host query front-end (parser + flattening) and the structural scoring rules vs the reference's exec_query — CPU only."""
import numpy as np
import pytest

import trinity_b200 as tb
from pyeval import evaluate
from refharness import RefIndex
from test_gpu_parity import TEMPLATES
from util import closed_form_lists

NDOCS = 30_000


@pytest.fixture(scope="module")
def small(ref):
    lists = closed_form_lists(NDOCS)
    r = RefIndex(ref, tb.CODEC_GOOGLE)
    names = [f"t{i + 1}" for i in range(len(lists))]
    for n, (d, f) in zip(names, lists):
        r.add_term(n, d, f)
    r.finish(NDOCS)
    return r, lists, tb.TermDictionary(names)


EXTRA = ["t1 OR t2 AND t3", "t1 AND t2 OR t3", "t1 OR t2 OR t3 AND t4 NOT t5", "(t1 AND t2 AND t3) OR t7",
         "t1 AND (t2 AND (t3 OR t4))", "(t1 OR (t2 OR t3)) AND t5", "t1 NOT t2 NOT t3", "t5 AND t1 NOT (t2 AND t3)",
         "(t1 NOT t2) OR (t3 NOT t5)", "t2 -t3 t5", "t7|t9|t10"]


@pytest.mark.parametrize("q", TEMPLATES + EXTRA)
def test_frontend_docs_and_scores_match_reference(small, q):
    r, lists, tdict = small
    nodes = tb.parse_query(q, tdict)
    for x in nodes:
        if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
            x["weight"] = tb.bm25_idf(len(lists[int(x["term"])][0]), NDOCS)
    m, s = evaluate(nodes, lists, NDOCS, weights=True)
    want, _ = r.exec(q, False, NDOCS + 1)
    assert np.array_equal(np.flatnonzero(m).astype(np.uint32), want), q
    if "nosuchterm" in q:
        return
    wd, ws = r.exec(q, True, NDOCS + 1)
    assert np.array_equal(np.flatnonzero(m).astype(np.uint32), wd)
    got = s[wd]
    rel = np.abs(got - ws) / np.maximum(np.abs(ws), 1e-30)
    assert rel.max() <= 1e-5, (q, int(rel.argmax()), got[rel.argmax()], ws[rel.argmax()])


def test_bm25_weight_matches_reference_scorer(small):
    r, lists, _ = small
    for t in range(len(lists)):
        idf = tb.bm25_idf(len(lists[t][0]), NDOCS)
        for freq in (0, 1, 2, 7, 63, 64, 300, 65535):
            a, b = tb.bm25_score(idf, freq), r.bm25(t, freq)
            assert abs(a - b) <= 1e-6 * max(abs(b), 1e-30), (t, freq, a, b)


def test_parse_errors():
    td = tb.TermDictionary(["a", "b"])
    for bad in ["", "(a AND b", "a AND", "AND a", "a )"]:
        with pytest.raises(tb.TrinityError):
            tb.parse_query(bad, td)


def test_abi_exports_every_declared_symbol():
    import ctypes, re
    from pathlib import Path
    from trinity_b200._ffi import EXPORTS, lib
    hdr = (Path(__file__).resolve().parent.parent / "include" / "trinity_b200.h").read_text()
    declared = set(re.findall(r"\b(trn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    L = lib()
    for s in EXPORTS:
        assert hasattr(L, s), s


OPTIONAL_QUERIES = ["t3 AND <t5>", "<t2> t7", "t3 AND t4 AND <t5>", "t3 <t5> <t7>", "<t5> t3 <t7> t4", "(t3 OR t4) <t5 OR t2>",
                    "t9 <t2 AND t3>", "(t3 <t5>) OR t7", "t3 <t5> NOT t2", "t3 <t5> <t7> <t2>"]


@pytest.mark.parametrize("q", OPTIONAL_QUERIES)
def test_optional_semantics_match_reference(small, q):
    """<expr> (ParseConstTrueExpr) next to a conjunction operand -> DocsSetIterators::Optional: main side decides the match,
    the optional side only adds its score where it is on the document"""
    r, lists, tdict = small
    nodes = tb.parse_query(q, tdict)
    assert tb.NODE_OPTIONAL in [int(k) for k in nodes["kind"]]
    for x in nodes:
        if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
            x["weight"] = tb.bm25_idf(len(lists[int(x["term"])][0]), NDOCS)
    m, s = evaluate(nodes, lists, NDOCS, weights=True)
    ids = np.flatnonzero(m).astype(np.uint32)
    want, _ = r.exec(q, False, NDOCS + 1, parser_flags=8)
    assert np.array_equal(ids, want)
    wd, ws = r.exec(q, True, NDOCS + 1, parser_flags=8)
    assert np.array_equal(ids, wd)
    rel = np.abs(s[wd] - ws) / np.maximum(np.abs(ws), 1e-30)
    assert rel.max() <= 1e-5


SOME_QUERIES = [("[t1, t2, t3]", 2), ("[t1, t2, t3, t4, t5]", 3), ("[t1, t2, t3]", 1), ("[t1, t2, t3]", 3), ("[t1, t2]", 3),
                ("[t3, t4 AND t5, t6 OR t7, t2]", 2), ("t1 AND [t2, t3, t4]", 2), ("[t2, t3, t4] NOT t1", 2), ("[t2, t3, t4] OR t9", 2),
                ("[t2, nosuchterm, t4, t5]", 2), ("[t1, t2, t3, t4, t5, t6, t7, t8, t9, t10]", 5), ("(t1 AND t2) OR [t3, t4, t5]", 2), ("[t3 t4, t5, t6]", 2)]
# (a MatchSome group nested inside another one sends the reference's compiler into a runaway allocation: not tested)


@pytest.mark.parametrize("q,m", SOME_QUERIES)
def test_match_some_semantics_match_reference(small, q, m):
    """[a, b, ...] with match_some.min = m -> DocsSetIterators::DisjunctionSome (docset_iterators.cpp:679-811): the documents at least m
    children match, scored with the sum of the children that match"""
    r, lists, tdict = small
    nodes = tb.parse_query(q, tdict, min_match=m)
    for x in nodes:
        if x["kind"] == tb.NODE_TERM and x["term"] != tb.EMPTY_TERM:
            x["weight"] = tb.bm25_idf(len(lists[int(x["term"])][0]), NDOCS)
    mm, s = evaluate(nodes, lists, NDOCS, weights=True)
    ids = np.flatnonzero(mm).astype(np.uint32)
    want, _ = r.exec(q, False, NDOCS + 1, parser_flags=16, min_match=m)
    assert np.array_equal(ids, want), (len(ids), len(want))
    wd, ws = r.exec(q, True, NDOCS + 1, parser_flags=16, min_match=m)
    assert np.array_equal(ids, wd)
    if len(wd):
        rel = np.abs(s[wd] - ws) / np.maximum(np.abs(ws), 1e-30)
        assert rel.max() <= 1e-5


def test_truth_tables_of_the_candidate_planner_match_the_evaluator(small):
    """the host planner of the candidate-driven path tabulates a tree's boolean function over its terms: every assignment must agree with
    the structural evaluator (pyeval) on one-document-per-assignment posting lists, and `necessary` must be the terms all matches hold"""
    _, _, tdict = small
    qs = [(q, None) for q in TEMPLATES + EXTRA + OPTIONAL_QUERIES] + list(SOME_QUERIES)
    checked = 0
    for q, m in qs:
        nodes = tb.parse_query(q, tdict, min_match=m)
        try:
            terms, table, necessary = tb.query_truth_table(nodes)
        except tb.TrinityError:
            continue
        n = len(terms)
        # document a+1 holds exactly the terms whose bit is set in assignment a
        ndocs = 1 << n
        lists = {t: (np.array([a + 1 for a in range(ndocs) if (a >> j) & 1], np.uint32), None) for j, t in enumerate(terms)}
        full = [(np.zeros(0, np.uint32), np.zeros(0, np.uint32))] * len(tdict)
        for t, (d, _) in lists.items():
            full[t] = (d, np.ones(len(d), np.uint32))
        # the evaluator mirrors the reference's root-filter quirk, and so does the planner (it tabulates the effective root): compare
        # on the plain tree semantics by evaluating node 0 without the quirk
        from pyeval import evaluate
        mm, _ = evaluate(nodes, full, ndocs, weights=None, quirk=False)
        want = np.array([mm[a + 1] for a in range(ndocs)], bool)
        assert np.array_equal(table, want), q
        nec = (1 << n) - 1
        for a in range(ndocs):
            if want[a]:
                nec &= a
        assert necessary == nec, q
        checked += 1
    assert checked >= 30


def test_parser_rejects_hostile_input_without_crashing(small):
    """query text comes from users: random token soup, 200000-deep nesting and million-term chains must end in a parse error (or a plan)"""
    _, _, tdict = small
    rng = np.random.default_rng(1)
    toks = ["t1", "t2", "t3", "t9", "nosuch", "AND", "OR", "NOT", "(", ")", "[", "]", ",", "<", ">", "-", "|", "||", " ", "  "]
    ok = bad = 0
    for _ in range(5000):
        s = "".join(str(rng.choice(toks)) + (" " if rng.random() < 0.7 else "") for _ in range(int(rng.integers(1, 25))))
        try:
            nodes = tb.parse_query(s, tdict)
            assert len(nodes) >= 1
            ok += 1
        except tb.TrinityError:
            bad += 1
    assert ok > 50 and bad > 50
    for o, c in (("(", ")"), ("[", "]"), ("<", ">")):
        with pytest.raises(tb.TrinityError):
            tb.parse_query(o * 200_000 + "t1" + c * 200_000, tdict)
        assert len(tb.parse_query(o * 50 + "t1" + c * 50, tdict)) == 1
    for op in (" AND ", " OR ", " NOT ", " "):
        with pytest.raises(tb.TrinityError):
            tb.parse_query(op.join(["t1", "t2", "t3"] * 100_000), tdict)


def test_truth_table_rejects_malformed_plans():
    """plans arrive through the C ABI: cycles, children in front of their parent and 100-level chains are argument errors"""
    from trinity_b200._ffi import QNODE_DTYPE
    chain = np.zeros(101, QNODE_DTYPE)
    for i in range(100):  # AND -> AND -> ... -> term
        chain[i] = (tb.NODE_AND, 1, i + 1, 0, 0.0)
    chain[100] = (tb.NODE_TERM, 0, 0, 3, 0.0)
    with pytest.raises(tb.TrinityError):
        tb.query_truth_table(chain)
    cyc = np.zeros(2, QNODE_DTYPE)
    cyc[0] = (tb.NODE_AND, 1, 1, 0, 0.0)
    cyc[1] = (tb.NODE_OR, 1, 0, 0, 0.0)  # points back at its parent
    with pytest.raises(tb.TrinityError):
        tb.query_truth_table(cyc)
    okp = np.zeros(3, QNODE_DTYPE)
    okp[0] = (tb.NODE_AND, 2, 1, 0, 0.0)
    okp[1] = (tb.NODE_TERM, 0, 0, 4, 0.0)
    okp[2] = (tb.NODE_TERM, 0, 0, 7, 0.0)
    terms, table, nec = tb.query_truth_table(okp)
    assert sorted(terms) == [4, 7] and table.tolist() == [False, False, False, True] and nec == 3
