"""GPU-side Encoder (trn_encode_google == Codecs::Google::Encoder, google_codec.cpp:9-176): the index built on the device
is BYTE-IDENTICAL to the one the reference's own encoder writes for the same postings (oracle/_ref through RefIndex.add_term), term tuples
included — blocks of 1 / 32 / 33 documents, terms without documents, freq-0 documents, 1..5-byte varbyte codes in deltas, freqs and position
deltas, hundreds of hits per document, more than SKIPLIST_STEP blocks (skiplist entries) and the countdown that carries over between terms.
The non-reference geometries of the decode sweep (block size / skiplist step) are checked against this repo's host encoder, and the encoded
index is executed: same matches as the reference's exec_query on its own index."""
import numpy as np
import pytest

import trinity_b200 as tb
from refharness import RefIndex
from util import assert_same_docs

pytestmark = pytest.mark.gpu


MAX_POSITION = 1 << 14  # Trinity::Limits::MaxPosition (trinity_limits.h:15): the reference encoder asserts pos < MaxPosition


def _list(rng, n, max_gap, freq_of, span=MAX_POSITION):
    docs = np.cumsum(rng.integers(1, max_gap, n, dtype=np.uint64)).astype(np.uint32) if n else np.zeros(0, np.uint32)
    freqs = np.array([freq_of(i) for i in range(n)], np.uint32)
    # sorted random positions in [1, span): repeated positions are legal (non-decreasing), deltas << 1 take 1, 2 and 3-byte codes
    pos = [np.sort(rng.integers(1, span, int(f))) for f in freqs]
    pos = np.concatenate(pos).astype(np.uint32) if pos else np.zeros(0, np.uint32)
    return docs, freqs, pos


def _shapes(rng):
    return [
        ("one-doc", _list(rng, 1, 50, lambda i: 2)),
        ("exactly-32", _list(rng, 32, 50, lambda i: 1 + i % 3)),
        ("empty", _list(rng, 0, 50, lambda i: 0)),
        ("33", _list(rng, 33, 50, lambda i: i % 2)),                                  # freq-0 documents
        ("skiplist", _list(rng, 32 * 21 + 5, 300, lambda i: 1 + (i * 7) % 4)),          # 22 blocks: entries, countdown carried in and out
        ("wide-gaps", _list(rng, 700, 3_000_000, lambda i: 1)),                         # 3/4-byte docID deltas, 2/3-byte position deltas
        ("huge-freq", _list(rng, 40, 20, lambda i: 130 + 200 * (i % 5), span=3000)),  # 2-byte freqs, blocks of several KB
        ("empty-again", _list(rng, 0, 50, lambda i: 0)),
        ("dense", _list(rng, 5000, 2, lambda i: 1 + (i % 17 == 0), span=40)),
    ]


def _five_byte_codes():
    # docID deltas >= 2^28 (5-byte codes), freq >= 2^14 (3-byte code; its 17 000 hits share the 16 383 legal positions)
    rng = np.random.default_rng(3)
    docs = np.array([7, 7 + (1 << 28) + 3, 7 + (1 << 29), 4_000_000_000], np.uint32)
    freqs = np.array([1, 2, 17_000, 1], np.uint32)
    pos = [np.array([MAX_POSITION - 1]), np.array([5, 9000]), np.sort(rng.integers(1, MAX_POSITION, 17_000)), np.array([3])]
    return docs, freqs, np.concatenate(pos).astype(np.uint32)


def test_device_encoder_is_byte_identical_to_the_reference_encoder(ref):
    rng = np.random.default_rng(77)
    shapes = _shapes(rng) + [("5-byte", _five_byte_codes())]
    r = RefIndex(ref, tb.CODEC_GOOGLE)
    for name, (d, f, p) in shapes:
        r.add_term(name, d, f, p)
    r.finish(int(max(int(d.max()) if d.size else 0 for _, (d, f, p) in shapes)))
    g = tb.GpuIndexSource(0)
    index, terms, countdown, ms = g.encode_google([l for _, l in shapes])
    want, wterms = r.index(), r.terms()
    assert index.size == want.size
    assert np.array_equal(index, want), f"first differing byte at {int(np.flatnonzero(index != want)[0])}"
    assert np.array_equal(terms, wterms)
    assert ms > 0
    nblocks = sum((len(d) + 31) // 32 for _, (d, f, p) in shapes)
    assert countdown == 8 - nblocks % 8
    g.close()


@pytest.mark.parametrize("block_docs,step,countdown", [(32, 8, 3), (8, 1, 1), (16, 64, 64), (128, 8, 8), (100, 3, 2), (1, 8, 5)])
def test_device_encoder_geometries_match_the_host_encoder(block_docs, step, countdown):
    """block size / skiplist step of the decode sweep + a session whose countdown is mid-way (terms encoded after others)"""
    rng = np.random.default_rng(block_docs * 131 + step)
    lists = [l for _, l in _shapes(rng)]
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    b.set_google_block(block_docs, step)
    # the host builder starts a session with a full countdown: a mid-way session = `step - countdown` filler blocks encoded first
    filler = step - countdown
    fl = [(np.arange(1, filler * block_docs + 1, dtype=np.uint32), np.ones(filler * block_docs, np.uint32), None)] if filler else []
    enc_countdown = countdown
    for d, f, p in fl:
        b.add_term(d, f, None)
    skip = b.index().size
    for d, f, p in lists:
        b.add_term(d, f, p)
    want = b.index()[skip:]
    wterms = b.terms_array()[len(fl):].copy()
    wterms["chunk_off"] -= skip
    g = tb.GpuIndexSource(0)
    index, terms, cd, _ = g.encode_google(lists, block_docs, step, enc_countdown)
    assert index.size == want.size
    assert np.array_equal(index, want), f"first differing byte at {int(np.flatnonzero(index != want)[0])}"
    assert np.array_equal(terms, wterms)
    g.close()


def test_device_encoder_without_positions_and_bad_input():
    rng = np.random.default_rng(5)
    lists = [(d, f, None) for _, (d, f, p) in _shapes(rng)]
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    for d, f, _ in lists:
        b.add_term(d, f, None)
    g = tb.GpuIndexSource(0)
    index, terms, _, _ = g.encode_google(lists)
    assert np.array_equal(index, b.index()) and np.array_equal(terms, b.terms_array())
    # what the reference encoder throws on is refused, loudly
    d = np.array([5, 9, 9, 12], np.uint32)
    with pytest.raises(tb.TrinityError, match="rc=-1"):
        g.encode_google([(d, np.ones(4, np.uint32), None)])
    with pytest.raises(tb.TrinityError, match="rc=-1"):
        g.encode_google([(np.array([0, 3], np.uint32), np.ones(2, np.uint32), None)])
    with pytest.raises(tb.TrinityError, match="rc=-1"):
        g.encode_google([(np.array([3, 4], np.uint32), np.array([2, 1], np.uint32), np.array([7, 5, 1], np.uint32))])
    with pytest.raises(tb.TrinityError, match="rc=-1"):  # Limits::MaxPosition
        g.encode_google([(np.array([3], np.uint32), np.array([1], np.uint32), np.array([MAX_POSITION], np.uint32))])
    g.close()


def test_an_index_encoded_on_the_device_executes_like_the_reference(ref):
    """encode on the GPU -> upload -> exec: the consumer of the bytes is the kernel set of the hot path, the judge the reference's exec_query
    over the index ITS encoder wrote from the same postings"""
    ndocs, nterms = 300_000, 24
    lists, names = [], []
    for rank in range(1, nterms + 1):
        d, f = tb.SynthIndex.postings(ndocs, rank, 500, 11)
        p = tb.SynthIndex.positions(ndocs, rank, 500, 11)
        lists.append((d, f, p))
        names.append(f"t{rank:04d}")
    r = RefIndex(ref, tb.CODEC_GOOGLE)
    for n, (d, f, p) in zip(names, lists):
        r.add_term(n, d, f, p)
    r.finish(ndocs)
    g = tb.GpuIndexSource(0)
    index, terms, _, _ = g.encode_google(lists)
    assert np.array_equal(index, r.index())
    g.upload(tb.CODEC_GOOGLE, index, terms, ndocs)
    tdict = tb.TermDictionary(names)
    qs = ["t0001 AND t0002", "t0003 OR t0017 OR t0024", "t0002 NOT t0005", "(t0001 OR t0009) AND (t0004 OR t0020) NOT t0003", '"t0001 t0002"']
    res = g.exec_batch([tb.parse_query(q, tdict) for q in qs], tb.MODE_DOCS_ONLY)
    for i, q in enumerate(qs):
        assert_same_docs(res.query(i)[0], r.exec(q, False, ndocs + 1)[0], q)
    g.close()
