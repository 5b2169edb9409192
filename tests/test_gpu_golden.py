"""GPU path vs the committed golden vectors (no oracle at run time): decode of reference-authored bytes and exec_query digests."""
import numpy as np
import pytest

import trinity_b200 as tb
from golden_util import check_docs_digest, check_scores_digest, load_closed, load_lists
from util import closed_form_lists

pytestmark = pytest.mark.gpu
CODECS = [tb.CODEC_GOOGLE, tb.CODEC_LUCENE]


@pytest.mark.parametrize("codec", CODECS)
def test_gpu_decodes_reference_authored_bytes(codec):
    z, n = load_lists(codec)
    g = tb.GpuIndexSource(0)
    g.upload(codec, z["index"], z["terms"], int(max(int(z[f"docids_{i}"][-1]) for i in range(n))))
    d, f, sums, _ = g.decode_terms(range(n), materialise=True)
    at = 0
    for i in range(n):
        m = len(z[f"docids_{i}"])
        assert np.array_equal(d[at:at + m], z[f"docids_{i}"]), f"term {i} docids"
        assert np.array_equal(f[at:at + m] & 0xFFFF, z[f"freqs_{i}"]), f"term {i} freqs"
        at += m


@pytest.mark.parametrize("codec", CODECS)
def test_gpu_exec_reproduces_golden_results(codec):
    z = load_closed(codec)
    ndocs = int(z["ndocs"][0])
    lists = closed_form_lists(ndocs)
    b = tb.IndexBuilder(codec)
    for d, f in lists:
        b.add_term(d, f)
    g = tb.GpuIndexSource(0)
    g.upload(codec, b.index(), b.terms_array(), ndocs)
    tdict = tb.TermDictionary([f"t{i + 1}" for i in range(len(lists))])
    qs = z["queries"].tolist()
    plans = [g.set_bm25_weights(tb.parse_query(q, tdict), ndocs) for q in qs]
    res = g.exec_batch(plans, tb.MODE_DOCS_ONLY)
    for qi, q in enumerate(qs):
        check_docs_digest(z, qi, res.query(qi)[0], f"[{q}]")
    scored = [qi for qi in range(len(qs)) if f"ssum_{qi}" in z]
    res = g.exec_batch([plans[qi] for qi in scored], tb.MODE_SCORED_ALL)
    for j, qi in enumerate(scored):
        ids, sc = res.query(j)
        check_docs_digest(z, qi, ids, f"[{qs[qi]}] scored")
        check_scores_digest(z, qi, ids, sc, f"[{qs[qi]}]")
