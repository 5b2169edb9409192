"""BASELINE.json's full size (100M documents, 4096 Zipfian terms) through size-independent checks: per-query match counts and docID
checksums of the GPU batch == the reference's exec_query on the same index bytes (oracle/_ref as the checker), for the headline
2-term AND batch and the 8-term trees; OR/BM25 top-100 on the Lucene codec: counts and top-k scores within 1e-5."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import trinity_b200 as tb
from refharness import RefIndex

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pytestmark = pytest.mark.gpu
N, V = 100_000_000, 4096


def _setup(ref, codec):
    threads = os.cpu_count() or 8
    synth = tb.SynthIndex(codec, N, V, threads=threads)
    g = tb.GpuIndexSource(0)
    g.upload(synth.codec, np.asarray(synth.index), np.asarray(synth.terms), N)
    r = RefIndex.from_bytes(ref, synth.codec, np.asarray(synth.index), np.asarray(synth.hits), synth.names, np.asarray(synth.terms), N, synth.sum_hits)
    return synth, g, r, tb.TermDictionary(synth.names), threads


def _checksums(res, n):
    off = np.asarray(res.offsets[: n + 1], np.int64)
    cs = np.concatenate([np.zeros(1, np.uint64), np.cumsum(np.asarray(res.docids[: off[-1]], np.uint64), dtype=np.uint64)])  # keep uint64
    return cs[off[1:]] - cs[off[:-1]]


def test_full_size_set_queries_match_reference(ref):
    from bench import gen_queries
    synth, g, r, tdict, threads = _setup(ref, tb.CODEC_GOOGLE)
    for workload, nq in (("and2", 300), ("tree8", 120)):
        texts, _ = gen_queries(workload, nq, V)
        res = g.exec_batch([tb.parse_query(t, tdict) for t in texts], tb.MODE_DOCS_ONLY)
        _, counts, sums, _, _ = r.exec_batch(texts, False, 100, threads)
        bad = np.flatnonzero(np.asarray(res.match_counts, np.uint64) != counts)
        assert len(bad) == 0, f"{workload}: match counts differ for queries {bad[:8]}: gpu {res.match_counts[bad[:8]]} ref {counts[bad[:8]]}"
        assert np.array_equal(_checksums(res, nq), sums), f"{workload}: docID checksums differ"


def test_full_size_or_topk_matches_reference(ref):
    from bench import gen_queries
    synth, g, r, tdict, threads = _setup(ref, tb.CODEC_LUCENE)
    nq, k = 48, 100
    texts, _ = gen_queries("or10", nq, V)
    plans = [g.set_bm25_weights(tb.parse_query(t, tdict), N) for t in texts]
    res = g.exec_batch(plans, tb.MODE_SCORED_TOPK, k)
    _, counts, _, tid, tsc = r.exec_batch(texts, True, k, threads)
    assert np.array_equal(np.asarray(res.match_counts, np.uint64), counts)
    for q in range(nq):
        d, s = res.query(q)
        assert len(s) == min(k, int(counts[q]))
        rel = np.abs(np.asarray(s, np.float64) - tsc[q][: len(s)]) / np.maximum(np.abs(tsc[q][: len(s)]), 1e-30)
        assert rel.max() <= 1e-5, f"query {q}: top-k score mismatch {rel.max():.3e}"
