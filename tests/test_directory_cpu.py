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
