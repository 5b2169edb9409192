"""Load-time block directory vs the decoded lists (CPU-only)."""
import numpy as np
import pytest

import trinity_b200 as tb
from test_codecs_cpu import make_lists, positions_for


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_directory_matches_postings(codec):
    rng = np.random.default_rng(7 + codec)
    lists = make_lists(rng)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f, positions_for(f, rng))
    index = b.index()
    bs = 32 if codec == tb.CODEC_GOOGLE else 128
    for (d, f), t in zip(lists, b.terms):
        last, off, first = tb.directory_probe(codec, index, t)
        nblocks = (len(d) + bs - 1) // bs
        assert len(last) == nblocks + 1 and last[-1] == 0xFFFFFFFF
        assert first == d[0]
        want = [d[min(len(d), (i + 1) * bs) - 1] for i in range(nblocks)]
        assert np.array_equal(last[:-1], np.array(want, np.uint32))
        assert np.all(np.diff(off.astype(np.int64)) > 0)
        assert t[1] <= off[0] and off[-1] <= t[1] + t[2]


def test_directory_lucene_long_list_uses_and_checks_skiplist():
    # > 2 full blocks so that the skiplist fast path (entries i>0 with a successor) is exercised
    rng = np.random.default_rng(5)
    d = np.cumsum(rng.integers(1, 300, 128 * 40 + 5)).astype(np.uint32)
    f = rng.integers(1, 5, len(d)).astype(np.uint32)
    b = tb.IndexBuilder(tb.CODEC_LUCENE)
    t = b.add_term(d, f)
    last, off, first = tb.directory_probe(tb.CODEC_LUCENE, b.index(), t)
    assert len(last) == 42 and first == d[0]
    assert np.array_equal(last[:40], d[127::128][:40]) and last[40] == d[-1]


def test_directory_rejects_corrupt_chunk():
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    t = b.add_term(np.arange(1, 200, dtype=np.uint32), np.ones(199, np.uint32))
    bad = b.index().copy()
    bad[4] ^= 0x7F  # block length byte of the first block
    with pytest.raises(tb.TrinityError):
        tb.directory_probe(tb.CODEC_GOOGLE, bad, t)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_directory_build_survives_corrupted_chunks(codec):
    """index bytes come from files: a corrupted or truncated term chunk must be reported (TRN_ERR_FORMAT) or parsed into an in-bounds
    directory — never read outside the buffer (that would kill the process)"""
    rng = np.random.default_rng(11 + codec)
    lists = make_lists(rng)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f, positions_for(f, rng))
    good = b.index().copy()
    ok = bad = 0
    for trial in range(400):
        idx = good.copy()
        t = b.terms[trial % len(b.terms)]
        off, ln = int(t[1]), int(t[2])
        for _ in range(int(rng.integers(1, 6))):
            idx[off + int(rng.integers(0, ln))] = int(rng.integers(0, 256))
        term = t
        if trial % 5 == 0:  # a chunk that claims more bytes / documents than it has
            term = (int(t[0]) + int(rng.integers(0, 500)), off, min(ln + int(rng.integers(0, 64)), idx.size - off))
        try:
            last, offs, first = tb.directory_probe(codec, idx, term)
            assert np.all(offs[:-1] >= off) and np.all(offs <= off + int(term[2]) + 8)
            ok += 1
        except tb.TrinityError:
            bad += 1
    assert ok + bad == 400


# ---------------------------------------------------------------------------------------------- sparse docID -> block tables
def _brute_first_block_ge(last, d):
    """first block whose last docID >= d (nblocks if none) == what skiplist_search + header hops of Decoder::advance arrive at"""
    return np.searchsorted(last, d, side="left").astype(np.uint32)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_sparse_table_lookup_equals_brute_force(codec):
    """the kernels' own lookup code (csrc/dirlookup.h, run on the host through trn_directory_lookup) against a plain searchsorted, on
    dense, sparse, clustered, tiny and shard-like (first docID far from 1) lists; probes sit on block ends, table boundaries, 0 and 2^32-1"""
    rng = np.random.default_rng(11 + codec)
    bs = 32 if codec == tb.CODEC_GOOGLE else 128
    shapes = {
        "dense": np.cumsum(rng.integers(1, 3, 60_000)),
        "sparse": np.cumsum(rng.integers(1, 40_000, 3_000)),
        "clustered": np.concatenate([np.arange(1, 5_000), 10_000_000 + np.cumsum(rng.integers(1, 9, 20_000)), [4_000_000_000, 4_000_000_123]]),
        "shard": 75_000_000 + np.cumsum(rng.integers(1, 50, 40_000)),
        "tiny": np.array([7]),
        "few": np.cumsum(rng.integers(1, 100_000, 5 * bs)),
        "huge-gaps": np.cumsum(rng.integers(1, 1 << 20, 4_000)),
    }
    for name, d in shapes.items():
        d = d.astype(np.uint32)
        b = tb.IndexBuilder(codec)
        t = b.add_term(d, np.ones(len(d), np.uint32))
        index = b.index()
        last = tb.directory_probe(codec, index, t)[0][:-1]
        nblocks = len(last)
        probes = np.concatenate([last, last + 1, last - 1, d[:: max(1, len(d) // 500)], rng.integers(0, 1 << 32, 2000, dtype=np.uint64).astype(np.uint32),
                                 (np.arange(0, 1 << 19) << 13).astype(np.uint32)[:: 37], [0, 1, d[0], d[0] - 1, d[-1], d[-1] + 1, 0xFFFFFFFF]]).astype(np.uint32)
        got, shift, entries = tb.directory_lookup(codec, index, t, probes)
        want = _brute_first_block_ge(last, probes)
        bad = np.flatnonzero(got != want)
        assert len(bad) == 0, f"{name}: docID {probes[bad[0]]} -> block {got[bad[0]]}, want {want[bad[0]]} (tf_shift {shift})"
        if nblocks <= 8:
            assert shift == 32 and entries == 0
        else:  # at most one table entry per two blocks (+ the closing entry), never finer than 512 docIDs
            assert 9 <= shift <= 31 and entries <= nblocks // 2 + 1, (name, shift, entries, nblocks)


@pytest.mark.parametrize("codec", [tb.CODEC_GOOGLE, tb.CODEC_LUCENE])
def test_directory_is_linear_in_blocks_and_terms_also_per_shard(codec):
    """the whole directory stays a small fraction of the index it describes, for the full synthetic index and for every docID-range
    shard of it (a shard's tables cover only the shard's own range) — the dense term x tile table it replaces was 24x the index here"""
    from trinity_b200.sharded import shard_range
    ndocs, nterms = 4_000_000, 512
    for world in (1, 8):
        for rank in sorted({0, world - 1}):
            s = tb.SynthIndex(codec, ndocs, nterms, min_df=50, threads=4, doc_range=shard_range(ndocs, rank, world))
            st = tb.directory_stats(codec, np.asarray(s.index), np.asarray(s.terms), threads=4)
            assert st["table_entries"] <= st["total_blocks"] + nterms
            assert st["directory_bytes"] <= 8 * (st["total_blocks"] + nterms) + 4 * st["table_entries"] + 36 * nterms
            assert st["directory_bytes"] <= 0.15 * st["index_bytes"], (world, rank, st)


def test_directory_of_a_million_tiny_terms_stays_small():
    """10^6 terms of 1..3 postings each (a real vocabulary's tail): per-term cost is the 36-byte record + its two block entries,
    no table at all — the dense table would have needed terms x tiles x 4 bytes (tens of GB at 100M documents)"""
    nterms = 1_000_000
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    rng = np.random.default_rng(2)
    base = rng.integers(1, 90_000_000, nterms)
    L = b._L
    import ctypes as C
    from trinity_b200._ffi import TrnTerm
    t = TrnTerm()
    terms = np.zeros(nterms, tb._ffi.TERM_DTYPE)
    one = np.ones(3, np.uint32)
    for i in range(nterms):
        n = 1 + (i % 3)
        d = (int(base[i]) + np.arange(n, dtype=np.uint32) * 1000).astype(np.uint32)
        L.trn_builder_add_term(b._h, d.ctypes.data_as(C.c_void_p), one.ctypes.data_as(C.c_void_p), n, None, C.byref(t))
        terms[i] = (t.documents, t.chunk_off, t.chunk_len)
    st = tb.directory_stats(tb.CODEC_GOOGLE, b.index(), terms, threads=4)
    assert st["total_blocks"] == nterms and st["table_entries"] == 0
    assert st["directory_bytes"] == nterms * (36 + 2 * 8)


@pytest.mark.parametrize("block_docs,step", [(8, 8), (16, 1), (64, 4), (128, 64)])
def test_directory_follows_the_block_size_of_sweep_indexes(block_docs, step):
    """GOOGLE-layout indexes built with other block sizes / skiplist steps (the decode sweep of BASELINE.json configs[4]): the directory
    finds the block size in the bytes; the reference format's (32, 8) stays the default"""
    rng = np.random.default_rng(block_docs)
    d = np.cumsum(rng.integers(1, 300, block_docs * 11 + 5)).astype(np.uint32)
    b = tb.IndexBuilder(tb.CODEC_GOOGLE)
    b.set_google_block(block_docs, step)
    t = b.add_term(d, np.ones(len(d), np.uint32))
    last, off, first = tb.directory_probe(tb.CODEC_GOOGLE, b.index(), t)
    assert len(last) == 12 + 1 and first == d[0]
    assert np.array_equal(last[:11], d[block_docs - 1::block_docs][:11]) and last[11] == d[-1]
    entries = int(np.frombuffer(b.index()[t[1]:t[1] + 2].tobytes(), "<u2")[0])
    assert entries == 12 // step  # one skiplist entry per `step` committed blocks
