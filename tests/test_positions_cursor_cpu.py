"""The kernels' position cursors (csrc/hitcursor.h: the per-(candidate, term) walk of phrase.cuh) run on the HOST through
trn_debug_positions and checked against the corpus, both codecs: GOOGLE inline hits (google_codec.cpp:497-594) and LUCENE hits.data through
the load-time hits directory (lucene_codec.cpp:401-513, :767-856) — documents at block starts / ends, in the varbyte tails, runs of hits
that cross 128-hit blocks, documents the term does not hold, PFor pages with exceptions."""
import numpy as np
import pytest

import trinity_b200 as tb


def _term(rng, ndocs_in_term, max_gap, freq_of):
    docs = np.cumsum(rng.integers(1, max_gap, ndocs_in_term)).astype(np.uint32)
    freqs = np.array([freq_of(i) for i in range(ndocs_in_term)], np.uint32)
    positions = []
    for f in freqs:
        gaps = rng.integers(1, 9, int(f))
        big = rng.random(int(f)) < 0.03  # a few large position gaps: exceptions in the PFor pages of the deltas
        gaps = np.where(big, gaps + rng.integers(5_000, 70_000, int(f)), gaps)
        positions.append(np.cumsum(gaps).astype(np.uint32))
    return docs, freqs, positions


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE], ids=["google", "lucene"])
def test_cursor_reads_the_positions_the_corpus_holds(codec):
    rng = np.random.default_rng(31 + codec)
    shapes = {
        "many-small": _term(rng, 1000, 50, lambda i: 1 + (i % 3)),                    # 3 full LUCENE blocks + tail; hits cross blocks everywhere
        "few-huge": _term(rng, 40, 5000, lambda i: 150 + 37 * (i % 5)),              # every document spans hit blocks
        "tail-only": _term(rng, 77, 9, lambda i: 1 + (i % 2)),                       # no full document block; < 128 hits in all -> the hits tail
        "exact-blocks": _term(rng, 256, 300, lambda i: 1),                           # 256 hits: two full hit blocks, empty tail
        "mixed": _term(rng, 700, 2000, lambda i: 1 if i % 7 else 200),
        "one": _term(rng, 1, 9, lambda i: 5),
    }
    b = tb.IndexBuilder(codec)
    terms = {}
    for name, (docs, freqs, positions) in shapes.items():
        terms[name] = b.add_term(docs, freqs, np.concatenate(positions))
    index, hits = b.index(), b.hits()
    for name, (docs, freqs, positions) in shapes.items():
        pick = np.unique(np.concatenate([[0, len(docs) - 1], rng.integers(0, len(docs), 60),
                                         np.arange(0, len(docs), 127)[:20], np.arange(127, len(docs), 128)[:8], np.arange(128, len(docs), 128)[:8]]))
        pick = pick[pick < len(docs)]
        absent = np.setdiff1d(np.concatenate([docs[pick] + 1, [1, int(docs[-1]) + 5]]).astype(np.uint32), docs)
        probe = np.concatenate([docs[pick], absent]).astype(np.uint32)
        got = tb.debug_positions(codec, index, hits, terms[name], probe)
        for j, i in enumerate(pick):
            assert np.array_equal(got[j], positions[i]), (name, int(i), got[j][:8], positions[i][:8])
        for j in range(len(pick), len(probe)):
            assert len(got[j]) == 0, (name, int(probe[j]))


def test_hits_directory_rejects_a_truncated_stream():
    rng = np.random.default_rng(2)
    docs, freqs, positions = _term(rng, 600, 50, lambda i: 2)
    b = tb.IndexBuilder(tb.CODEC_LUCENE)
    t = b.add_term(docs, freqs, np.concatenate(positions))
    index, hits = b.index(), b.hits()
    with pytest.raises(tb.TrinityError):
        tb.debug_positions(tb.CODEC_LUCENE, index, hits[: len(hits) // 2], t, docs[-3:])
