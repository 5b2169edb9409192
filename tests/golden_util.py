"""Loading + checking helpers for the committed golden fixtures (tests/golden/, generated from the reference by make_golden.py)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
CODEC_NAME = {0: "google", 1: "lucene"}


def load_lists(codec):
    z = np.load(GOLDEN / f"lists_{CODEC_NAME[codec]}.npz")
    n = int(z["nlists"][0])
    return z, n


def load_closed(codec):
    return np.load(GOLDEN / f"closed_form_{CODEC_NAME[codec]}.npz")


def load_widened(codec):
    return np.load(GOLDEN / f"widened_{CODEC_NAME[codec]}.npz")


def check_docs_digest(z, qi, ids, what):
    ids = np.asarray(ids, np.uint32)
    assert len(ids) == int(z[f"count_{qi}"][0]), f"{what}: count {len(ids)} != golden {int(z[f'count_{qi}'][0])}"
    assert int(ids.astype(np.uint64).sum()) == int(z[f"sum_{qi}"][0]), f"{what}: docID sum differs"
    assert int(np.bitwise_xor.reduce(ids) if len(ids) else 0) == int(z[f"xor_{qi}"][0]), f"{what}: docID xor differs"
    assert np.array_equal(ids[:16], z[f"head_{qi}"]) and np.array_equal(ids[-16:], z[f"tail_{qi}"]), f"{what}: head/tail differ"
    assert np.all(np.diff(ids.astype(np.int64)) > 0), f"{what}: not strictly ascending"


def check_scores_digest(z, qi, ids, scores, what, rtol=1e-5):
    ids, scores = np.asarray(ids, np.uint32), np.asarray(scores, np.float64)
    assert abs(scores.sum() - float(z[f"ssum_{qi}"][0])) <= rtol * max(1.0, abs(float(z[f"ssum_{qi}"][0]))), f"{what}: score sum"
    order = np.lexsort((ids, -scores))[:16]
    gs, ws = scores[order], z[f"tops_{qi}"]
    assert np.all(np.abs(gs - ws) <= rtol * np.abs(ws) + 1e-12), f"{what}: top-16 scores differ {gs} vs {ws}"
    # docIDs of the top-16 must agree wherever the golden scores are not tied (within tolerance) with a neighbour
    wd = z[f"topd_{qi}"]
    for i in range(len(wd)):
        tied = any(abs(ws[i] - ws[j]) <= 4 * rtol * abs(ws[i]) for j in range(len(wd)) if j != i) or i == len(wd) - 1
        if not tied:
            assert ids[order][i] == wd[i], f"{what}: top doc #{i} {ids[order][i]} != {wd[i]}"
