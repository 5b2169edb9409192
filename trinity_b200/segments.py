"""A search session over several Trinity segments == IndexSourcesCollection (index_source.h:191-238, index_source.cpp:3-30).

Sources are ordered newest generation first; segment i is scanned with the updated_documents of every NEWER segment as its
masked-documents registry (fused into the emission stage on the device, trn_set_masked_documents); BM25 statistics are the
collection's (Σ docsCnt, Σ document frequency over the sources — similarity.h:202-222).  Each segment's postings live in HBM in
their own engine context; a query batch runs once per segment, exactly like the per-source exec_query() loop of a Trinity
application."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

from . import (EMPTY_TERM, NODE_TERM, GpuIndexSource, Segment, TermDictionary, bm25_idf, parse_query)


def generation_of(path: str) -> int:
    """a segment directory's name is its generation (segment_index_source.cpp:18-21)"""
    return int(os.path.basename(os.path.normpath(str(path))))


class SegmentCollection:
    def __init__(self, paths: Sequence[str], device: int = 0, max_docid: int | None = None):
        paths = sorted((str(p) for p in paths), key=generation_of, reverse=True)
        self.generations = [generation_of(p) for p in paths]
        if len(set(self.generations)) != len(paths):
            raise ValueError("no two sources may share a generation")
        self.segments: List[Segment] = [Segment(p) for p in paths]
        self.dicts = [TermDictionary(s.names) for s in self.segments]
        self.docs_cnt = sum(s.field_statistics["docsCnt"] for s in self.segments)
        self._df = {}
        for s in self.segments:
            for n, t in zip(s.names, s.terms):
                self._df[n] = self._df.get(n, 0) + int(t["documents"])
        self.sources: List[GpuIndexSource] = []
        newer = np.zeros(0, np.uint32)
        for s in self.segments:
            g = GpuIndexSource(device)
            # the directory does not record the largest docID: 0 = the block directory built at upload finds it
            g.upload(s.codec, s.index, s.terms, max_docid or 0)
            if newer.size:
                g.set_masked_documents(newer)
            self.sources.append(g)
            if s.masked_documents.size:
                newer = np.union1d(newer, s.masked_documents).astype(np.uint32)

    def document_frequency(self, term: str) -> int:
        return self._df.get(term, 0)

    def plans(self, text: str, scored: bool):
        """one plan per segment (term ids are per dictionary); BM25 weights from the collection's statistics"""
        out = []
        for s, d in zip(self.segments, self.dicts):
            nodes = parse_query(text, d)
            if scored:
                for x in nodes:
                    if x["kind"] == NODE_TERM and x["term"] != EMPTY_TERM:
                        x["weight"] = bm25_idf(self._df[s.names[int(x["term"])]], self.docs_cnt)
            out.append(nodes)
        return out

    def exec_batch(self, queries: Sequence[str], mode: int, k: int = 100):
        """-> [BatchResult per segment], collection order (newest first)"""
        from . import MODE_DOCS_ONLY
        scored = mode != MODE_DOCS_ONLY
        per_q = [self.plans(q, scored) for q in queries]
        return [g.exec_batch([pq[i] for pq in per_q], mode, k) for i, g in enumerate(self.sources)]
