// Segment-directory ingestion: reads the files a Trinity SegmentIndexSession::commit() writes
// (indexer.cpp:241-300 persist_segment, terms.cpp:125-170 pack_terms, docidupdates.cpp:8-72 pack_updates) so that the engine runs on
// indexes produced by Trinity's own indexer — the host half of SegmentIndexSource (segment_index_source.cpp:5-186).
//
//   <dir>/index                  raw postings chunks (uploaded to HBM unchanged by trn_upload_index)
//   <dir>/terms.data             { u8 commonPrefixLen, u8 suffixLen, suffix, varuint32 documents, varuint32 chunkLen, u32 chunkOffset }*
//   <dir>/id                     u8 1, u8 len, codec id ("GOOGLE" / "LUCENE"), u64 sumTermHits, u32 totalTerms, u64 sumTermsDocs, u32 docsCnt
//   <dir>/updated_documents.ids  banks of 32K-doc bitmaps [+ 32 KB bloom filter], u8 log2(bank), u8 noBloom, u32 bankBase[], u32 nBanks, u32 lo, u32 hi
// (terms.idx is only a skiplist over terms.data; hits.data holds Lucene positions, which this path never reads.)
#include "../../include/trinity_b200.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

struct trn_segment {
        int                       codec{0};
        std::vector<uint8_t>      index;
        std::vector<trn_term>     terms;
        std::vector<std::string>  names;
        std::vector<const char *> namePtrs;
        std::vector<uint32_t>     masked;
        uint64_t                  sumTermHits{0}, sumTermsDocs{0};
        uint32_t                  totalTerms{0}, docsCnt{0};
};

namespace {
bool read_file(const std::string &path, std::vector<uint8_t> &out, bool required) {
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) {
                if (required)
                        throw std::runtime_error("cannot open " + path);
                return false;
        }
        const std::streamsize n = f.tellg();
        f.seekg(0);
        out.resize(size_t(n));
        if (n && !f.read(reinterpret_cast<char *>(out.data()), n))
                throw std::runtime_error("cannot read " + path);
        return true;
}

// LEB128 as written by Compression::PackUInt32 (Switch/compress.h:65-106)
uint32_t varuint32(const uint8_t *&p, const uint8_t *e) {
        uint32_t v{0};
        for (uint32_t shift = 0; shift < 35; shift += 7) {
                if (p >= e)
                        throw std::runtime_error("terms.data: truncated varuint");
                const uint8_t b = *p++;
                v |= uint32_t(b & 0x7fu) << shift;
                if (b < 128)
                        return v;
        }
        throw std::runtime_error("terms.data: malformed varuint");
}
uint32_t rd32(const uint8_t *p) {
        return p[0] | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
uint64_t rd64(const uint8_t *p) {
        return uint64_t(rd32(p)) | (uint64_t(rd32(p + 4)) << 32);
}
} // namespace

extern "C" int trn_segment_open(const char *dir, trn_segment **out, char *err, size_t errcap) {
        if (!dir || !out)
                return TRN_ERR_ARG;
        auto seg = new trn_segment();
        try {
                const std::string    base(dir);
                std::vector<uint8_t> buf;
                // ---- id: codec + default field statistics
                read_file(base + "/id", buf, true);
                if (buf.size() < 2 || buf[0] != 1 || buf.size() < size_t(2 + buf[1] + 24))
                        throw std::runtime_error("id: unsupported release or truncated");
                const std::string codec(reinterpret_cast<const char *>(buf.data() + 2), buf[1]);
                if (codec == "GOOGLE")
                        seg->codec = TRN_CODEC_GOOGLE;
                else if (codec == "LUCENE")
                        seg->codec = TRN_CODEC_LUCENE;
                else
                        throw std::runtime_error("id: unknown codec '" + codec + "'");
                const uint8_t *p  = buf.data() + 2 + buf[1];
                seg->sumTermHits  = rd64(p);
                seg->totalTerms   = rd32(p + 8);
                seg->sumTermsDocs = rd64(p + 12);
                seg->docsCnt      = rd32(p + 20);
                // ---- index
                read_file(base + "/index", seg->index, true);
                // ---- terms.data: front-coded dictionary, every term carries its term_index_ctx
                if (read_file(base + "/terms.data", buf, false)) {
                        const uint8_t *q = buf.data(), *const e = q + buf.size();
                        std::string    prev;
                        while (q < e) {
                                if (q + 2 > e)
                                        throw std::runtime_error("terms.data: truncated entry");
                                const uint32_t common = q[0], suffix = q[1];
                                q += 2;
                                if (common > prev.size() || q + suffix > e)
                                        throw std::runtime_error("terms.data: bad prefix/suffix lengths");
                                std::string term = prev.substr(0, common) + std::string(reinterpret_cast<const char *>(q), suffix);
                                q += suffix;
                                trn_term t;
                                t.documents = varuint32(q, e);
                                t.chunk_len = varuint32(q, e);
                                if (q + 4 > e)
                                        throw std::runtime_error("terms.data: truncated chunk offset");
                                t.chunk_off = rd32(q);
                                q += 4;
                                if (uint64_t(t.chunk_off) + t.chunk_len > seg->index.size())
                                        throw std::runtime_error("terms.data: chunk of '" + term + "' exceeds the index file");
                                seg->terms.push_back(t);
                                seg->names.push_back(term);
                                prev.swap(term);
                        }
                }
                seg->namePtrs.reserve(seg->names.size());
                for (auto &n : seg->names)
                        seg->namePtrs.push_back(n.c_str());
                // ---- updated_documents.ids: the docIDs this (newer) segment masks in OLDER segments
                if (read_file(base + "/updated_documents.ids", buf, false) && !buf.empty()) {
                        if (buf.size() < 14)
                                throw std::runtime_error("updated_documents.ids: truncated trailer");
                        const uint8_t *e      = buf.data() + buf.size();
                        const uint32_t nbanks = rd32(e - 12);
                        const uint8_t *skip   = e - 12 - size_t(nbanks) * 4;
                        if (skip - 2 < buf.data())
                                throw std::runtime_error("updated_documents.ids: bad skiplist size");
                        const uint32_t bankBits = 1u << skip[-2];
                        const bool     noBloom  = skip[-1] != 0;
                        const size_t   bankBytes = bankBits / 8;
                        const size_t   need      = size_t(nbanks) * bankBytes + (noBloom ? 0 : 256 * 1024 / 8);
                        if (size_t(skip - 2 - buf.data()) != need)
                                throw std::runtime_error("updated_documents.ids: size does not match its trailer");
                        for (uint32_t b = 0; b < nbanks; ++b) {
                                const uint32_t bankBase = rd32(skip + size_t(b) * 4);
                                const uint8_t *bm       = buf.data() + size_t(b) * bankBytes;
                                for (uint32_t i = 0; i < bankBits; ++i)
                                        if (bm[i >> 3] & (1u << (i & 7)))
                                                seg->masked.push_back(bankBase + i);
                        }
                }
                *out = seg;
                return TRN_OK;
        } catch (const std::exception &ex) {
                if (err && errcap) {
                        std::strncpy(err, ex.what(), errcap - 1);
                        err[errcap - 1] = 0;
                }
                delete seg;
                return TRN_ERR_FORMAT;
        }
}

extern "C" void trn_segment_close(trn_segment *s) {
        delete s;
}

extern "C" int trn_segment_info(trn_segment *s, int *codec, uint32_t *nterms, uint64_t *index_bytes, uint64_t *sum_term_hits, uint32_t *total_terms,
                                uint64_t *sum_terms_docs, uint32_t *docs_cnt, uint64_t *nmasked) {
        if (!s)
                return TRN_ERR_ARG;
        if (codec) *codec = s->codec;
        if (nterms) *nterms = uint32_t(s->terms.size());
        if (index_bytes) *index_bytes = s->index.size();
        if (sum_term_hits) *sum_term_hits = s->sumTermHits;
        if (total_terms) *total_terms = s->totalTerms;
        if (sum_terms_docs) *sum_terms_docs = s->sumTermsDocs;
        if (docs_cnt) *docs_cnt = s->docsCnt;
        if (nmasked) *nmasked = s->masked.size();
        return TRN_OK;
}

extern "C" int trn_segment_index(trn_segment *s, const uint8_t **index, uint64_t *nbytes) {
        if (!s || !index || !nbytes)
                return TRN_ERR_ARG;
        *index  = s->index.data();
        *nbytes = s->index.size();
        return TRN_OK;
}

extern "C" int trn_segment_terms(trn_segment *s, const trn_term **terms, const char *const **names, uint32_t *nterms) {
        if (!s || !terms || !names || !nterms)
                return TRN_ERR_ARG;
        *terms  = s->terms.data();
        *names  = s->namePtrs.data();
        *nterms = uint32_t(s->terms.size());
        return TRN_OK;
}

extern "C" int trn_segment_masked(trn_segment *s, const uint32_t **docids, uint64_t *n) {
        if (!s || !docids || !n)
                return TRN_ERR_ARG;
        *docids = s->masked.data();
        *n      = s->masked.size();
        return TRN_OK;
}
