"""Host write path vs the reference encoders: byte-exact index (+hits.data) for both codecs, and the reference
decoders reading OUR bytes.  CPU-only."""
import numpy as np
import pytest

import trinity_b200 as tb
from refharness import RefIndex


def make_lists(rng, n_lists=12):
    """hand-built lists hitting the format edge cases SURVEY.md 8c lists"""
    out = []
    sizes = [1, 2, 31, 32, 33, 127, 128, 129, 256, 1000, 128 * 9 + 77, 5000]
    for i, n in enumerate(sizes[:n_lists]):
        style = i % 4
        if style == 0:      # dense, 1-byte gaps, constant freq (all-equal freq blocks)
            gaps = rng.integers(1, 4, n)
            freqs = np.full(n, 1)
        elif style == 1:    # mixed gaps incl. 2/3-byte varbytes, varied freqs
            gaps = rng.integers(1, 40000, n)
            freqs = rng.integers(1, 9, n)
        elif style == 2:    # mostly small with rare huge gaps (PFor exceptions, 4-byte varbytes)
            gaps = np.where(rng.random(n) < 0.05, rng.integers(1 << 21, 1 << 22, n), rng.integers(1, 16, n))
            freqs = np.where(rng.random(n) < 0.1, rng.integers(20, 140, n), rng.integers(1, 3, n))
        else:               # constant gap (all-equal delta blocks) + some zero freqs (docs without hits)
            gaps = np.full(n, 7)
            freqs = rng.integers(0, 3, n)
        docids = np.cumsum(gaps).astype(np.uint32)
        out.append((docids, freqs.astype(np.uint32)))
    # a 5-byte varbyte delta (>= 2^28) and maxbits-b == 1 exceptions
    g = np.ones(300, dtype=np.int64)
    g[5] = 3 * 10 ** 8
    g[130:258:3] = 2     # 1-bit exceptions over b=1
    out.append((np.cumsum(g).astype(np.uint32), np.ones(300, np.uint32)))
    return out


def positions_for(freqs, rng):
    pos = []
    for f in freqs:
        p = 0
        for _ in range(int(f)):
            p += int(rng.integers(1, 18))
            pos.append(p)
    return np.array(pos, dtype=np.uint32)


def pfor_padding_mask(index_ref, index_mine):
    """The reference's FastPFor leaves the 0-3 padding bytes after the exception-positions byte array uninitialised
    (fastpfor.h:196-198 memcpy without clearing) — those bytes carry no information.  Everything else must match."""
    return np.flatnonzero(index_ref != index_mine)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_encoder_bytes_match_reference(ref, codec):
    rng = np.random.default_rng(1234 + codec)
    lists = make_lists(rng)
    r = RefIndex(ref, codec)
    b = tb.IndexBuilder(codec)
    for i, (d, f) in enumerate(lists):
        pos = positions_for(f, rng)
        r.add_term(f"t{i}", d, f, pos)
        b.add_term(d, f, pos)
    r.finish(int(max(int(d[-1]) for d, _ in lists)))
    mine, theirs = b.index(), r.index()
    assert np.array_equal(b.terms_array(), r.terms()), "term_index_ctx (documents, offset, size) differ"
    assert mine.size == theirs.size
    diff = np.flatnonzero(mine != theirs)
    if codec == tb.CODEC_GOOGLE:
        assert diff.size == 0, f"google index differs at {diff[:10]}"
    else:
        # tolerate ONLY uninitialised PFor padding bytes of the reference: our bytes there are 0
        assert np.all(mine[diff] == 0), f"lucene index differs at non-padding bytes {diff[:10]}"
        assert diff.size < mine.size // 50
        hm, ht = b.hits(), r.hits()
        assert hm.size == ht.size
        hd = np.flatnonzero(hm != ht)
        assert np.all(hm[hd] == 0) and hd.size < max(1, hm.size // 50)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_reference_decoder_reads_our_bytes(ref, codec):
    rng = np.random.default_rng(99 + codec)
    lists = make_lists(rng)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f, positions_for(f, rng))
    names = [f"t{i}" for i in range(len(lists))]
    r = RefIndex.from_bytes(ref, codec, b.index(), b.hits(), names, b.terms_array(), int(max(int(d[-1]) for d, _ in lists)))
    for i, (d, f) in enumerate(lists):
        dd, ff = r.decode(i, len(d) + 8)
        assert np.array_equal(dd, d)
        assert np.array_equal(ff, f & 0xFFFF)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_synth_index_equals_reference_encoding_of_same_postings(ref, codec):
    ndocs, nterms, min_df = 200_000, 24, 50
    s = tb.SynthIndex(codec, ndocs, nterms, min_df=min_df, threads=3)
    r = RefIndex(ref, codec)
    total = 0
    for rank in range(1, nterms + 1):
        d, f = tb.SynthIndex.postings(ndocs, rank, min_df)
        p = tb.SynthIndex.positions(ndocs, rank, min_df)
        assert len(d) == max(min_df, ndocs // (2 * rank)) and d[-1] <= ndocs and np.all(np.diff(d.astype(np.int64)) > 0)
        assert f.min() >= 1 and f.max() <= 8 and len(p) == int(f.sum())
        r.add_term(s.names[rank - 1], d, f, p)
        total += int(f.sum())
    r.finish(ndocs)
    assert s.sum_hits == total
    assert np.array_equal(np.asarray(s.terms), r.terms())
    mine, theirs = np.asarray(s.index), r.index()
    diff = np.flatnonzero(mine != theirs)
    if codec == tb.CODEC_GOOGLE:
        assert diff.size == 0
    else:
        assert np.all(mine[diff] == 0) and diff.size < mine.size // 50


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_synth_docid_shards_partition_the_index(ref, codec):
    """docID-range shards (multi-GPU layout): every shard is byte-identical to the reference encoding of the
    postings that fall into its range, and the shards together hold every posting exactly once."""
    ndocs, nterms, min_df, G = 120_000, 16, 40, 3
    bounds = [(g * ndocs // G + 1, (g + 1) * ndocs // G) for g in range(G)]
    counts = np.zeros(nterms, np.int64)
    for lo, hi in bounds:
        s = tb.SynthIndex(codec, ndocs, nterms, min_df=min_df, threads=2, doc_range=(lo, hi))
        r = RefIndex(ref, codec)
        for rank in range(1, nterms + 1):
            d, f = tb.SynthIndex.postings(ndocs, rank, min_df)
            p = tb.SynthIndex.positions(ndocs, rank, min_df)
            keep = (d >= lo) & (d <= hi)
            ends = np.cumsum(f)
            starts = ends - f
            pk = np.concatenate([p[a:b] for a, b, k in zip(starts, ends, keep) if k]) if keep.any() else np.zeros(0, np.uint32)
            r.add_term(s.names[rank - 1], d[keep], f[keep], pk)
            counts[rank - 1] += int(keep.sum())
        r.finish(ndocs)
        assert np.array_equal(np.asarray(s.terms), r.terms())
        mine, theirs = np.asarray(s.index), r.index()
        diff = np.flatnonzero(mine != theirs)
        assert np.all(mine[diff] == 0) and (codec == tb.CODEC_LUCENE or diff.size == 0)
    for rank in range(1, nterms + 1):
        assert counts[rank - 1] == max(min_df, ndocs // (2 * rank))


def test_reference_authored_synthetic_index_is_the_same_workload(ref):
    """bench.py's --impl reference arm authors its index with the reference's own Encoders (tref_synth_build: one IndexSession per term,
    in parallel) so that it never loads the product library.  GOOGLE: byte-identical to the product builder's index (term tuples too).
    LUCENE: identical term tuples (sizes, offsets) and identical decoded postings; the bytes may differ in don't-care padding bits of the
    FastPFor pages (a fresh reference Encoder packs from uninitialised scratch buffers)."""
    from refharness import RefIndex
    ndocs, nterms, min_df = 300_000, 48, 70
    for codec in (tb.CODEC_GOOGLE, tb.CODEC_LUCENE):
        r = RefIndex.synth_build(ref, codec, ndocs, nterms, min_df=min_df, threads=3)
        s = tb.SynthIndex(codec, ndocs, nterms, min_df=min_df, threads=2)
        assert np.array_equal(r.terms(), np.asarray(s.terms))
        if codec == tb.CODEC_GOOGLE:
            assert np.array_equal(r.index(), np.asarray(s.index))
        assert len(r.hits()) == len(np.asarray(s.hits))
        for ti in (0, 1, 7, nterms - 1):
            d, f = tb.SynthIndex.postings(ndocs, ti + 1, min_df)
            rd, rf = r.decode(ti, len(d) + 4)
            assert np.array_equal(rd, d) and np.array_equal(rf, f)
        # and it executes: the reference's exec_query over its own index == over the product builder's bytes
        r2 = RefIndex.from_bytes(ref, codec, np.asarray(s.index), np.asarray(s.hits), s.names, np.asarray(s.terms), ndocs, s.sum_hits)
        for q in ("t0001 AND t0002", "t0003 OR t0040 OR t0011"):
            a, _ = r.exec(q, False, ndocs + 1)
            b, _ = r2.exec(q, False, ndocs + 1)
            assert np.array_equal(a, b)
