"""ctypes wrapper of oracle/libtrinity_oracle.so (the plain-C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "oracle" / "libtrinity_oracle.so"


def load():
    if not SO.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "-s"])
    L = C.CDLL(str(SO))
    vp = C.c_void_p
    L.orc_decode_google.restype = C.c_int64
    L.orc_decode_google.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint64]
    L.orc_decode_lucene.restype = C.c_int64
    L.orc_decode_lucene.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64]
    L.orc_bm25_idf.restype = C.c_double
    L.orc_bm25_idf.argtypes = [C.c_uint32, C.c_uint64]
    L.orc_bm25_score.restype = C.c_float
    L.orc_bm25_score.argtypes = [C.c_double, C.c_uint16]
    L.orc_positions_google.restype = C.c_int64
    L.orc_positions_google.argtypes = [vp, C.c_uint32, vp, C.c_uint64]
    L.orc_positions_lucene.restype = C.c_int64
    L.orc_positions_lucene.argtypes = [vp, C.c_uint32, vp, vp, C.c_uint32, vp, C.c_uint64]
    L.orc_exec2.restype = C.c_int
    L.orc_exec2.argtypes = [C.c_int, vp, vp, vp, vp, C.c_uint32, C.c_int, vp, vp]
    L.orc_exec.restype = C.c_int
    L.orc_exec.argtypes = [C.c_int, vp, vp, vp, C.c_uint32, C.c_int, vp, vp]
    return L


def decode(L, codec, index, term):
    docs, off, ln = int(term[0]), int(term[1]), int(term[2])
    d, f = np.zeros(docs + 1, np.uint32), np.zeros(docs + 1, np.uint32)
    chunk = np.ascontiguousarray(index[off:off + ln])
    if codec == 0:
        n = L.orc_decode_google(chunk.ctypes.data, ln, d.ctypes.data, f.ctypes.data, docs)
    else:
        n = L.orc_decode_lucene(chunk.ctypes.data, ln, docs, d.ctypes.data, f.ctypes.data, docs)
    assert n == docs, (n, docs)
    return d[:docs], f[:docs]


def positions(L, codec, index, hits, term):
    """flat positions (freq entries per document, list order) of one term, restated decode of the inline hits / hits.data"""
    docs, off, ln = int(term[0]), int(term[1]), int(term[2])
    chunk = np.ascontiguousarray(index[off:off + ln])
    _, f = decode(L, codec, index, term)
    cap = int(f.astype(np.uint64).sum())
    out = np.zeros(max(cap, 1), np.uint32)
    if codec == 0:
        n = L.orc_positions_google(chunk.ctypes.data, ln, out.ctypes.data, cap)
    else:
        hits = np.ascontiguousarray(hits, np.uint8)
        f = np.ascontiguousarray(f, np.uint32)
        n = L.orc_positions_lucene(chunk.ctypes.data, ln, hits.ctypes.data, f.ctypes.data, docs, out.ctypes.data, cap)
    assert n == cap, (n, cap)
    return out[:cap]


def exec_query(L, codec, index, terms, nodes, ndocs, scored, hits=None):
    index = np.ascontiguousarray(index, np.uint8)
    terms = np.ascontiguousarray(terms)
    nodes = np.ascontiguousarray(nodes)
    m = np.zeros(ndocs + 1, np.uint8)
    s = np.zeros(ndocs + 1, np.float64)
    hits = None if hits is None or not len(hits) else np.ascontiguousarray(hits, np.uint8)
    rc = L.orc_exec2(codec, index.ctypes.data, None if hits is None else hits.ctypes.data, terms.ctypes.data, nodes.ctypes.data, ndocs, int(scored),
                     m.ctypes.data, s.ctypes.data)
    assert rc == 0
    ids = np.flatnonzero(m).astype(np.uint32)
    return ids, s[ids]
