#!/usr/bin/env python
"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF (oracle/_ref/libtrinity_ref.so = the
reference's own encoders, decoders and exec_query compiled in place).  Run in the build container (needs /root/reference
to have built oracle/_ref); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py

Outputs (small, committed):
  lists_{google,lucene}.npz   reference-encoded index bytes (+hits.data) of the hand-built edge-case lists, their term_index_ctx
                               and the (docID, freq) streams the reference PostingsListIterator yields for them
  widened_{codec}.npz         the rows added after SURVEY 8a-e: MatchSome groups on the closed-form index (same digests as below) and
                               phrases on a document-major corpus: reference-encoded index + hits.data bytes, term_index_ctx, and the
                               digests of 14 phrase queries
  closed_form_{codec}.npz     reference exec_query results on the closed-form index (term i = multiples of PRIMES[i], 200k docs):
                               per query: match count, sum of docIDs, xor of docIDs, first/last 16 docIDs, and for the scored
                               run the top-16 (docID, score) by (score desc, docID asc) plus the sum of all scores
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import trinity_b200 as tb  # noqa: E402  (only for constants + query text list)
from refharness import RefIndex, load_ref  # noqa: E402
from test_codecs_cpu import make_lists, positions_for  # noqa: E402
from test_frontend_cpu import EXTRA, SOME_QUERIES  # noqa: E402
from test_phrase_cpu import NDOCS as PHRASE_NDOCS, QUERIES as PHRASE_QUERIES, VOCAB as PHRASE_VOCAB  # noqa: E402
from test_gpu_parity import TEMPLATES  # noqa: E402
from util import closed_form_lists  # noqa: E402

GOLDEN_NDOCS = 200_000
QUERIES = [q for q in TEMPLATES + EXTRA]


def main():
    ref = load_ref()
    for codec, cname in ((tb.CODEC_GOOGLE, "google"), (tb.CODEC_LUCENE, "lucene")):
        rng = np.random.default_rng(20260924 + codec)
        lists = make_lists(rng)
        r = RefIndex(ref, codec)
        pos_all = []
        for i, (d, f) in enumerate(lists):
            p = positions_for(f, rng)
            pos_all.append(p)
            r.add_term(f"t{i}", d, f, p)
        r.finish(int(max(int(d[-1]) for d, _ in lists)))
        out = {"index": r.index(), "hits": r.hits(), "terms": r.terms(), "nlists": np.array([len(lists)])}
        for i, (d, f) in enumerate(lists):
            dd, ff = r.decode(i, len(d) + 4)
            assert np.array_equal(dd, d)
            out[f"docids_{i}"], out[f"freqs_{i}"], out[f"infreqs_{i}"], out[f"positions_{i}"] = dd, ff, f, pos_all[i]
        np.savez_compressed(HERE / f"lists_{cname}.npz", **out)

        r = RefIndex(ref, codec)
        cl = closed_form_lists(GOLDEN_NDOCS)
        for i, (d, f) in enumerate(cl):
            r.add_term(f"t{i + 1}", d, f)
        r.finish(GOLDEN_NDOCS)
        res = {"queries": np.array(QUERIES), "ndocs": np.array([GOLDEN_NDOCS])}
        for qi, q in enumerate(QUERIES):
            ids, _ = r.exec(q, False, GOLDEN_NDOCS + 1)
            res[f"count_{qi}"] = np.array([len(ids)], np.uint64)
            res[f"sum_{qi}"] = np.array([ids.astype(np.uint64).sum()], np.uint64)
            res[f"xor_{qi}"] = np.array([np.bitwise_xor.reduce(ids) if len(ids) else 0], np.uint32)
            res[f"head_{qi}"], res[f"tail_{qi}"] = ids[:16], ids[-16:]
            if "nosuchterm" in q:
                continue
            sid, sc = r.exec(q, True, GOLDEN_NDOCS + 1)
            assert np.array_equal(sid, ids)
            order = np.lexsort((sid, -sc))[:16]
            res[f"topd_{qi}"], res[f"tops_{qi}"] = sid[order], sc[order]
            res[f"ssum_{qi}"] = np.array([sc.sum()])
        np.savez_compressed(HERE / f"closed_form_{cname}.npz", **res)

        # ---- widened rows: MatchSome on the same closed-form index, phrases on a document-major corpus
        wid = {"ndocs": np.array([GOLDEN_NDOCS]), "some_queries": np.array([q for q, _ in SOME_QUERIES]), "some_min": np.array([m for _, m in SOME_QUERIES])}

        def digest(out, key, ids, sid=None, sc=None):
            out[f"count_{key}"] = np.array([len(ids)], np.uint64)
            out[f"sum_{key}"] = np.array([ids.astype(np.uint64).sum()], np.uint64)
            out[f"xor_{key}"] = np.array([np.bitwise_xor.reduce(ids) if len(ids) else 0], np.uint32)
            out[f"head_{key}"], out[f"tail_{key}"] = ids[:16], ids[-16:]
            if sc is not None:
                order = np.lexsort((sid, -sc))[:16]
                out[f"topd_{key}"], out[f"tops_{key}"] = sid[order], sc[order]
                out[f"ssum_{key}"] = np.array([sc.sum()])

        for qi, (q, m) in enumerate(SOME_QUERIES):
            ids, _ = r.exec(q, False, GOLDEN_NDOCS + 1, parser_flags=16, min_match=m)
            sid, sc = r.exec(q, True, GOLDEN_NDOCS + 1, parser_flags=16, min_match=m)
            assert np.array_equal(sid, ids)
            digest(wid, f"s{qi}", ids, sid, sc)
        prng = np.random.default_rng(21)
        prob = 1.0 / np.arange(1, PHRASE_VOCAB + 1)
        prob /= prob.sum()
        per_term = [dict() for _ in range(PHRASE_VOCAB)]
        for d in range(1, PHRASE_NDOCS + 1):
            toks = prng.choice(PHRASE_VOCAB, size=int(prng.integers(3, 30)), p=prob)
            for pos, t in enumerate(toks, start=1):
                per_term[int(t)].setdefault(d, []).append(pos)
        pr = RefIndex(ref, codec)
        for t in range(PHRASE_VOCAB):
            docs = np.array(sorted(per_term[t]), np.uint32)
            fr = np.array([len(per_term[t][int(d)]) for d in docs], np.uint32)
            flat = np.array([p for d in docs for p in per_term[t][int(d)]], np.uint32)
            pr.add_term(f"w{t + 1}", docs, fr, flat)
        pr.finish(PHRASE_NDOCS)
        wid.update({"phrase_index": pr.index(), "phrase_hits": pr.hits(), "phrase_terms": pr.terms(), "phrase_ndocs": np.array([PHRASE_NDOCS]),
                    "phrase_queries": np.array(PHRASE_QUERIES)})
        for qi, q in enumerate(PHRASE_QUERIES):
            ids, _ = pr.exec(q, False, PHRASE_NDOCS + 1)
            sid, sc = pr.exec(q, True, PHRASE_NDOCS + 1)
            assert np.array_equal(sid, ids)
            digest(wid, f"p{qi}", ids, sid, sc)
        np.savez_compressed(HERE / f"widened_{cname}.npz", **wid)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
