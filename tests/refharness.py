"""ctypes wrapper of oracle/_ref/libtrinity_ref.so (the reference's own hot path compiled in place by
oracle/build_ref.sh).  TEST INFRASTRUCTURE ONLY — never imported by the product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_SO = ROOT / "oracle" / "_ref" / "libtrinity_ref.so"
REF_GPU_SO = ROOT / "oracle" / "_ref" / "libtrinity_ref_gpu.so"  # the same reference objects + the reference-side binding of libtrinity_b200.so


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefLib:
    def __init__(self, L):
        self.L = L
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.tref_new.restype = vp
        L.tref_new.argtypes = [C.c_int]
        L.tref_free.argtypes = [vp]
        L.tref_add_term.argtypes = [vp, C.c_char_p, vp, vp, u32, vp]
        L.tref_finish.argtypes = [vp, u64]
        L.tref_from_bytes.restype = vp
        L.tref_from_bytes.argtypes = [C.c_int, vp, u64, vp, u64, vp, vp, vp, vp, u32, u64, u64]
        L.tref_index_size.restype = u64
        L.tref_index_size.argtypes = [vp]
        L.tref_index_data.restype = vp
        L.tref_index_data.argtypes = [vp]
        L.tref_hits_size.restype = u64
        L.tref_hits_size.argtypes = [vp]
        L.tref_hits_data.restype = vp
        L.tref_hits_data.argtypes = [vp]
        L.tref_num_terms.restype = u32
        L.tref_num_terms.argtypes = [vp]
        L.tref_term.argtypes = [vp, u32, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.tref_decode.restype = C.c_int64
        L.tref_decode.argtypes = [vp, u32, vp, vp, u64]
        L.tref_advance.argtypes = [vp, u32, vp, u32, vp]
        L.tref_positions.restype = C.c_int64
        L.tref_positions.argtypes = [vp, u32, vp, u64]
        L.tref_bm25.restype = C.c_double
        L.tref_bm25.argtypes = [vp, u32, u32]
        L.tref_exec.restype = C.c_int64
        L.tref_exec.argtypes = [vp, C.c_char_p, C.c_int, vp, vp, u64]
        L.tref_exec2.restype = C.c_int64
        L.tref_exec2.argtypes = [vp, C.c_char_p, C.c_int, u32, vp, vp, u64]
        L.tref_exec3.restype = C.c_int64
        L.tref_exec3.argtypes = [vp, C.c_char_p, C.c_int, u32, u32, vp, vp, u64]
        L.tref_exec_masked.restype = C.c_int64
        L.tref_exec_masked.argtypes = [vp, C.c_char_p, C.c_int, vp, u32, vp, vp, u64]
        L.tref_exec_batch.restype = C.c_double
        L.tref_exec_batch.argtypes = [vp, vp, u32, C.c_int, u32, C.c_int, vp, vp, vp, vp]
        L.tref_synth_build.restype = vp
        L.tref_synth_build.argtypes = [C.c_int, u32, u32, u32, u64, C.c_int, C.c_int]
        L.tref_last_error.restype = C.c_char_p
        L.tref_segment_write.argtypes = [C.c_int, C.c_char_p, u32, vp, vp, vp, vp, vp, u32, u32]
        L.tref_segment_open.restype = vp
        L.tref_segment_open.argtypes = [C.c_char_p]
        L.tref_collection_open.restype = vp
        L.tref_collection_open.argtypes = [vp, u32]
        L.tref_collection_exec.restype = C.c_int64
        L.tref_collection_exec.argtypes = [vp, C.c_char_p, C.c_int, vp, vp, u64, vp]
        L.tref_resolve.argtypes = [vp, C.c_char_p, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.tref_field_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(u64), C.POINTER(u32)]

    def collection_open(self, paths):
        """IndexSourcesCollection over several segment directories (scanned newest generation first)"""
        enc = [str(p).encode() for p in paths]
        arr = (C.c_char_p * len(enc))(*enc)
        h = self.L.tref_collection_open(C.cast(arr, C.c_void_p), len(enc))
        if not h:
            raise RuntimeError(self.err())
        x = RefIndex(self, -1, C.c_void_p(h))
        x.nsources = len(enc)
        return x

    def segment_write(self, codec: int, path: str, lists: dict, erased=(), replace_below: int = 0):
        """Index `lists` ({term: (docids, freqs)}) with the reference's SegmentIndexSession and commit() the segment to `path`
        (the last path component must be a number: the segment's generation).  Documents with id < replace_below are committed with
        replace() — updates of documents an older segment holds — and, like `erased`, land in updated_documents.ids."""
        names = list(lists)
        enc = [n.encode() for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        counts = np.array([len(lists[n][0]) for n in names], np.uint32)
        d = np.ascontiguousarray(np.concatenate([np.asarray(lists[n][0], np.uint32) for n in names]) if names else np.zeros(0, np.uint32))
        f = np.ascontiguousarray(np.concatenate([np.asarray(lists[n][1], np.uint32) for n in names]) if names else np.zeros(0, np.uint32))
        er = np.ascontiguousarray(erased, np.uint32)
        if self.L.tref_segment_write(codec, str(path).encode(), len(enc), C.cast(arr, C.c_void_p), _p(counts), _p(d), _p(f), _p(er), len(er), replace_below) != 0:
            raise RuntimeError(self.err())

    def segment_open(self, path: str):
        """The reference's SegmentIndexSource over `path`, wrapped for exec()"""
        h = self.L.tref_segment_open(str(path).encode())
        if not h:
            raise RuntimeError(self.err())
        return RefIndex(self, -1, C.c_void_p(h))

    def err(self):
        return self.L.tref_last_error().decode()


class RefIndex:
    """An in-memory index authored by the reference Encoders (or wrapping foreign bytes) + reference exec."""

    def __init__(self, rl: RefLib, codec: int, handle=None):
        self.rl, self.codec = rl, codec
        self.h = C.c_void_p(rl.L.tref_new(codec)) if handle is None else handle
        self.names = []

    @classmethod
    def from_bytes(cls, rl, codec, index, hits, names, terms, docs_cnt, sum_hits=0):
        index = np.ascontiguousarray(index, np.uint8)
        hits = np.ascontiguousarray(hits, np.uint8) if hits is not None and len(hits) else None
        enc = [n.encode() for n in names]
        arr = (C.c_char_p * len(enc))(*enc)
        docs = np.ascontiguousarray(terms["documents"], np.uint32)
        off = np.ascontiguousarray(terms["chunk_off"], np.uint32)
        ln = np.ascontiguousarray(terms["chunk_len"], np.uint32)
        h = rl.L.tref_from_bytes(codec, _p(index), index.size, _p(hits), 0 if hits is None else hits.size,
                                 C.cast(arr, C.c_void_p), _p(docs), _p(off), _p(ln), len(enc), docs_cnt, sum_hits)
        if not h:
            raise RuntimeError(rl.err())
        x = cls(rl, codec, C.c_void_p(h))
        x.names = list(names)
        return x

    @classmethod
    def synth_build(cls, rl, codec, ndocs, nterms, min_df=1000, seed=0x5EED, with_hits=True, threads=1):
        """the BASELINE.md synthetic index authored by the reference's own Encoders (no product code involved)"""
        h = rl.L.tref_synth_build(codec, ndocs, nterms, min_df, seed, int(with_hits), threads)
        if not h:
            raise RuntimeError(rl.err())
        x = cls(rl, codec, C.c_void_p(h))
        x.names = [f"t{r:04d}" for r in range(1, nterms + 1)]
        return x

    def add_term(self, name, docids, freqs, positions=None):
        d = np.ascontiguousarray(docids, np.uint32)
        f = np.ascontiguousarray(freqs, np.uint32)
        p = None if positions is None else np.ascontiguousarray(positions, np.uint32)
        r = self.rl.L.tref_add_term(self.h, name.encode(), _p(d), _p(f), len(d), _p(p))
        if r < 0:
            raise RuntimeError(self.rl.err())
        self.names.append(name)
        return r

    def finish(self, docs_cnt):
        if self.rl.L.tref_finish(self.h, docs_cnt) != 0:
            raise RuntimeError(self.rl.err())

    def index(self):
        n = self.rl.L.tref_index_size(self.h)
        return np.ctypeslib.as_array(C.cast(self.rl.L.tref_index_data(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def hits(self):
        n = self.rl.L.tref_hits_size(self.h)
        if not n:
            return np.zeros(0, np.uint8)
        return np.ctypeslib.as_array(C.cast(self.rl.L.tref_hits_data(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()

    def terms(self):
        from trinity_b200._ffi import TERM_DTYPE
        n = self.rl.L.tref_num_terms(self.h)
        out = np.zeros(n, TERM_DTYPE)
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        for i in range(n):
            self.rl.L.tref_term(self.h, i, C.byref(a), C.byref(b), C.byref(c))
            out[i] = (a.value, b.value, c.value)
        return out

    def decode(self, term_idx, cap):
        d, f = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        n = self.rl.L.tref_decode(self.h, term_idx, _p(d), _p(f), cap)
        if n < 0:
            raise RuntimeError(self.rl.err())
        return d[:n], f[:n]

    def positions(self, term_idx, cap):
        """flat positions of every document of the term (freq entries per document), via the reference's materialize_hits"""
        p = np.zeros(max(cap, 1), np.uint32)
        n = self.rl.L.tref_positions(self.h, term_idx, _p(p), cap)
        if n < 0:
            raise RuntimeError(self.rl.err())
        return p[:n]

    def advance(self, term_idx, targets):
        t = np.ascontiguousarray(targets, np.uint32)
        o = np.zeros(len(t), np.uint32)
        if self.rl.L.tref_advance(self.h, term_idx, _p(t), len(t), _p(o)) != 0:
            raise RuntimeError(self.rl.err())
        return o

    def collection_exec(self, q: str, scored: bool, cap: int):
        """per-source exec_query over the collection -> [(ids, scores)] in collection order"""
        ids = np.zeros(max(cap, 1), np.uint32)
        sc = np.zeros(max(cap, 1), np.float64)
        cnt = np.zeros(self.nsources, np.uint64)
        n = self.rl.L.tref_collection_exec(self.h, q.encode(), 1 if scored else 0, _p(ids), _p(sc), cap, _p(cnt))
        if n < 0:
            raise RuntimeError(self.rl.err())
        out, at = [], 0
        for c in cnt:
            c = int(c)
            out.append((ids[at:at + c], sc[at:at + c] if scored else None))
            at += c
        return out

    def resolve(self, term: str):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        if self.rl.L.tref_resolve(self.h, term.encode(), C.byref(a), C.byref(b), C.byref(c)) != 0:
            raise RuntimeError(self.rl.err())
        return a.value, b.value, c.value

    def field_stats(self):
        a, b, c, d = C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_uint32()
        if self.rl.L.tref_field_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)) != 0:
            raise RuntimeError(self.rl.err())
        return {"sumTermHits": a.value, "totalTerms": b.value, "sumTermsDocs": c.value, "docsCnt": d.value}

    def bm25(self, term_idx, freq):
        return self.rl.L.tref_bm25(self.h, term_idx, freq)

    def exec(self, q: str, scored: bool, cap: int, parser_flags: int = 0, min_match: int = 0):
        """parser_flags: ast_parser::Flags; 8 = ParseConstTrueExpr (the <expr> syntax -> Optional), 16 = ParseMatchSomeExpr ([a, b, c]);
        min_match: match_some.min of every MatchSome group (the parser leaves it at 1)"""
        ids = np.zeros(max(cap, 1), np.uint32)
        sc = np.zeros(max(cap, 1), np.float64)
        n = self.rl.L.tref_exec3(self.h, q.encode(), 1 if scored else 0, parser_flags, min_match, _p(ids), _p(sc), cap)
        if n < 0:
            raise RuntimeError(self.rl.err())
        assert n <= cap, "reference produced more matches than the capacity given"
        return ids[:n], (sc[:n] if scored else None)

    def exec_masked(self, q: str, scored: bool, masked, cap: int):
        """exec_query with the reference's masked_documents_registry holding `masked` docIDs"""
        mk = np.ascontiguousarray(masked, np.uint32)
        ids = np.zeros(max(cap, 1), np.uint32)
        sc = np.zeros(max(cap, 1), np.float64)
        n = self.rl.L.tref_exec_masked(self.h, q.encode(), 1 if scored else 0, _p(mk), len(mk), _p(ids), _p(sc), cap)
        if n < 0:
            raise RuntimeError(self.rl.err())
        return ids[:n], (sc[:n] if scored else None)

    def exec_batch(self, queries, scored: bool, k: int, threads: int):
        enc = [q.encode() for q in queries]
        arr = (C.c_char_p * len(enc))(*enc)
        nq = len(enc)
        counts = np.zeros(nq, np.uint64)
        sums = np.zeros(nq, np.uint64)
        tid = np.zeros((nq, k), np.uint32)
        tsc = np.zeros((nq, k), np.float64)
        el = self.rl.L.tref_exec_batch(self.h, C.cast(arr, C.c_void_p), nq, 1 if scored else 0, k, threads, _p(counts), _p(sums), _p(tid), _p(tsc))
        if el < 0:
            raise RuntimeError(self.rl.err())
        return el, counts, sums, tid, tsc

    def __del__(self):
        try:
            self.rl.L.tref_free(self.h)
        except Exception:
            pass


_ref = None
_ref_gpu = None


def load_ref_gpu() -> RefLib:
    """the reference compiled with its exec_query() span site going through integration/gpu_exec.cpp (needs a CUDA device at run time)"""
    global _ref_gpu
    if _ref_gpu is None:
        if not REF_GPU_SO.exists():
            subprocess.check_call(["bash", str(ROOT / "oracle" / "build_ref.sh")])
        rl = RefLib(C.CDLL(str(REF_GPU_SO)))
        rl.L.tref_gpu_attach.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        rl.L.tref_gpu_detach.argtypes = [C.c_void_p]
        rl.L.tref_gpu_spans_executed.restype = C.c_uint64
        rl.L.tref_gpu_spans_executed.argtypes = [C.c_void_p]
        _ref_gpu = rl
    return _ref_gpu


def load_ref() -> RefLib:
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            subprocess.check_call(["bash", str(ROOT / "oracle" / "build_ref.sh")])
        _ref = RefLib(C.CDLL(str(REF_SO)))
    return _ref
