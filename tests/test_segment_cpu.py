"""Segment directories written by the reference's own indexer (SegmentIndexSession::commit, indexer.cpp:241-300) are read by the
product's host reader (trn_segment_open) into exactly what the reference's SegmentIndexSource (segment_index_source.cpp:5-186)
restores: terms dictionary, term_index_ctx per term, field statistics, codec, masked documents."""
import numpy as np
import pytest

import trinity_b200 as tb
from refharness import load_ref


ERASED = np.array([3, 17, 4000, 40000, 70001, 1 << 20], np.uint32)  # a session cannot erase a document it also indexes


def make_lists(seed, nterms=40, ndocs=6000):
    rng = np.random.default_rng(seed)
    lists = {}
    pool = np.setdiff1d(np.arange(1, ndocs + 1), ERASED)
    for t in range(nterms):
        df = int(rng.integers(1, ndocs // (1 + t % 7)))
        docs = np.sort(rng.choice(pool, size=df, replace=False)).astype(np.uint32)
        freqs = rng.integers(1, 6, size=df).astype(np.uint32)
        # mixed-length names exercise the front coding of terms.data
        name = ("t%d" % t) if t % 3 else ("term_prefix_%03d" % t)
        lists[name] = (docs, freqs)
    return lists


@pytest.mark.parametrize("codec", [0, 1])
def test_segment_reader_matches_reference(tmp_path, codec):
    ref = load_ref()
    lists = make_lists(7 + codec)
    erased = ERASED
    path = tmp_path / "100"
    path.mkdir()
    ref.segment_write(codec, path, lists, erased, replace_below=100)   # documents 1..99 are updates of an older segment's
    seg = tb.Segment(str(path))
    rseg = ref.segment_open(path)
    assert seg.codec == codec
    assert sorted(seg.names) == sorted(lists)
    assert seg.names == sorted(seg.names), "terms.data is in dictionary order"
    for name, t in zip(seg.names, seg.terms):
        assert (int(t["documents"]), int(t["chunk_off"]), int(t["chunk_len"])) == rseg.resolve(name), name
        assert int(t["documents"]) == len(lists[name][0])
    assert rseg.resolve("no-such-term")[0] == 0
    assert seg.field_statistics == rseg.field_stats()
    assert seg.field_statistics["totalTerms"] == len(lists)
    assert seg.field_statistics["sumTermsDocs"] == sum(len(v[0]) for v in lists.values())
    updated = np.unique(np.concatenate([d[d < 100] for d, _ in lists.values()]))
    assert np.array_equal(seg.masked_documents, np.union1d(erased, updated))
    assert seg.index.size == (path / "index").stat().st_size


def test_segment_without_updates_and_errors(tmp_path):
    ref = load_ref()
    path = tmp_path / "7"
    path.mkdir()
    ref.segment_write(0, path, {"a": ([1, 2, 3], [1, 1, 2]), "ab": ([2], [1])})
    seg = tb.Segment(str(path))
    assert seg.masked_documents.size == 0 and seg.names == ["a", "ab"]
    with pytest.raises(tb.TrinityError):
        tb.Segment(str(tmp_path / "nope"))
    (path / "id").write_bytes(b"\x02\x06GOOGLE" + b"\0" * 24)
    with pytest.raises(tb.TrinityError):
        tb.Segment(str(path))
