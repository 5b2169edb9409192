// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-ABI harness around the *reference's own* classes (compiled in place from /root/reference by
// oracle/build_ref.sh into oracle/_ref/libtrinity_ref.so).  It is the parity oracle and the CPU
// baseline ("cpu_baseline.kind" = "reference"):
//   * builds in-memory indexes through the reference Encoders      (codecs.h:176-200)
//   * decodes postings through the reference PostingsListIterator  (codecs.h:211-246)
//   * runs the reference Trinity::exec_query()                     (exec.h:50-52)
//   * exposes the reference BM25 scorer                            (similarity.h:165-255)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
#include "exec.h"
#include "google_codec.h"
#include "indexer.h"
#include "lucene_codec.h"
#include "segment_index_source.h"
#include "similarity.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace Trinity;

namespace {
        thread_local std::string g_err;

        struct MemSrc final : public IndexSource {
                Codecs::AccessProxy *                           ap{nullptr};
                std::unordered_map<std::string, term_index_ctx> terms;
                field_statistics                                fs;

                term_index_ctx resolve_term_ctx(const str8_t term) override {
                        auto it = terms.find(std::string(term.data(), term.size()));
                        if (it == terms.end())
                                return {};
                        return it->second;
                }

                Codecs::Decoder *new_postings_decoder(const str8_t, const term_index_ctx ctx) override {
                        return ap->new_decoder(ctx);
                }

                field_statistics default_field_stats() override {
                        return fs;
                }

                bool index_empty() const override {
                        return false;
                }
        };

        struct RefIndex {
                int                                   codec{0};
                std::unique_ptr<Codecs::IndexSession> sess;
                std::unique_ptr<Codecs::Encoder>      enc;
                std::vector<uint8_t>                  index, hits;
                std::vector<term_index_ctx>           tctx;
                std::vector<std::string>              names;
                std::unique_ptr<Codecs::AccessProxy>  ap;
                IndexSource *                         src{nullptr};
                std::unique_ptr<IndexSourcesCollection> col;
                uint64_t                              sumHits{0};

                ~RefIndex() {
                        col.reset(); // releases src
                }

                void open(uint64_t docsCnt) {
                        if (codec == 0)
                                ap.reset(new Codecs::Google::AccessProxy("/tmp", index.data()));
                        else
                                ap.reset(new Codecs::Lucene::AccessProxy("/tmp", index.data(), hits.empty() ? (const uint8_t *)"" : hits.data()));
                        auto ms = new MemSrc();
                        src     = ms;
                        ms->ap  = ap.get();
                        uint64_t sumDocs{0};
                        for (size_t i = 0; i < names.size(); ++i) {
                                ms->terms.emplace(names[i], tctx[i]);
                                sumDocs += tctx[i].documents;
                        }
                        ms->fs.docsCnt      = docsCnt;
                        ms->fs.sumTermsDocs = sumDocs;
                        ms->fs.totalTerms   = names.size();
                        ms->fs.sumTermHits  = sumHits;
                        col.reset(new IndexSourcesCollection());
                        col->insert(src); // Retain()
                        src->Release();   // collection now holds the only ref
                        col->commit();
                }
        };

        struct CollectSink final : public MatchedIndexDocumentsFilter {
                std::vector<uint32_t> ids;
                std::vector<double>   scores;
                bool                  wantScores{false};
                uint64_t              n{0}, cap{0};

                void consider(const docid_t id) override {
                        if (n < cap)
                                ids.push_back(id);
                        ++n;
                }
                void consider(const docid_t id, const double score) override {
                        if (n < cap) {
                                ids.push_back(id);
                                scores.push_back(score);
                        }
                        ++n;
                }
        };

        // top-k sink: score desc, docID asc (the tie rule this repo defines; the reference leaves top-k to the app, matches.h:139-186)
        struct TopKSink final : public MatchedIndexDocumentsFilter {
                struct E {
                        double   s;
                        uint32_t id;
                };
                struct Worse {
                        // priority_queue top = the WORST kept element
                        bool operator()(const E &a, const E &b) const noexcept {
                                if (a.s != b.s)
                                        return a.s > b.s;
                                return a.id < b.id;
                        }
                };
                std::priority_queue<E, std::vector<E>, Worse> pq;
                uint32_t                                    k{100};
                uint64_t                                    n{0};

                void consider(const docid_t id, const double score) override {
                        ++n;
                        if (pq.size() < k)
                                pq.push({score, uint32_t(id)});
                        else {
                                const auto &w = pq.top();
                                if (score > w.s || (score == w.s && id < w.id)) {
                                        pq.pop();
                                        pq.push({score, uint32_t(id)});
                                }
                        }
                }
                void consider(const docid_t) override {
                        ++n;
                }
        };

        struct IdsSink final : public MatchedIndexDocumentsFilter {
                std::vector<uint32_t> ids;
                uint64_t              sum{0};
                void                  consider(const docid_t id) override {
                        ids.push_back(id);
                        sum += id;
                }
        };

        template <typename F>
        int guarded(F &&f) {
                try {
                        f();
                        return 0;
                } catch (const std::exception &e) {
                        g_err = e.what();
                } catch (...) {
                        g_err = "unknown exception";
                }
                return -1;
        }
} // namespace

extern "C" {
const char *tref_last_error() {
        return g_err.c_str();
}

// codec: 0 = GOOGLE, 1 = LUCENE (FastPFor<4>, the reference's compile-time default lucene_codec.h:21-29)
void *tref_new(int codec) {
        auto x   = new RefIndex();
        x->codec = codec;
        if (codec == 0)
                x->sess.reset(new Codecs::Google::IndexSession("/tmp"));
        else
                x->sess.reset(new Codecs::Lucene::IndexSession("/tmp"));
        x->sess->begin();
        x->enc.reset(x->sess->new_encoder());
        return x;
}

void tref_free(void *h) {
        delete static_cast<RefIndex *>(h);
}

// positions: absolute positions per hit, concatenated over documents (sum(freqs) entries); may be null => 1..freq
int tref_add_term(void *h, const char *name, const uint32_t *docids, const uint32_t *freqs, uint32_t n, const uint32_t *positions) {
        auto x = static_cast<RefIndex *>(h);
        int  idx{-1};
        if (guarded([&] {
                    term_index_ctx t;
                    size_t         pi{0};
                    x->enc->begin_term();
                    for (uint32_t i = 0; i < n; ++i) {
                            x->enc->begin_document(docids[i]);
                            for (uint32_t k = 0; k < freqs[i]; ++k) {
                                    const uint32_t pos = positions ? positions[pi++] : k + 1;
                                    x->enc->new_hit(pos, {});
                            }
                            x->sumHits += freqs[i];
                            x->enc->end_document();
                    }
                    x->enc->end_term(&t);
                    idx = int(x->tctx.size());
                    x->tctx.push_back(t);
                    x->names.emplace_back(name);
            }))
                return -1;
        return idx;
}

int tref_finish(void *h, uint64_t docsCnt) {
        auto x = static_cast<RefIndex *>(h);
        return guarded([&] {
                x->index.assign((const uint8_t *)x->sess->indexOut.data(), (const uint8_t *)x->sess->indexOut.data() + x->sess->indexOut.size());
                if (x->codec == 1) {
                        // take the hits.data bytes straight from the session buffer (IndexSession::end() would persist them to basePath/hits.data)
                        auto ls = static_cast<Codecs::Lucene::IndexSession *>(x->sess.get());
                        x->hits.assign((const uint8_t *)ls->positionsOut.data(), (const uint8_t *)ls->positionsOut.data() + ls->positionsOut.size());
                }
                x->enc.reset();
                x->sess.reset();
                x->open(docsCnt);
        });
}

void *tref_from_bytes(int codec, const uint8_t *index, uint64_t n, const uint8_t *hits, uint64_t nh, const char *const *names, const uint32_t *docs,
                      const uint32_t *off, const uint32_t *len, uint32_t nterms, uint64_t docsCnt, uint64_t sumHits) {
        auto x   = new RefIndex();
        x->codec = codec;
        x->index.assign(index, index + n);
        if (hits && nh)
                x->hits.assign(hits, hits + nh);
        for (uint32_t i = 0; i < nterms; ++i) {
                x->names.emplace_back(names[i]);
                x->tctx.emplace_back(docs[i], range32_t{off[i], len[i]});
        }
        x->sumHits = sumHits;
        if (guarded([&] { x->open(docsCnt); })) {
                delete x;
                return nullptr;
        }
        return x;
}

uint64_t tref_index_size(void *h) {
        return static_cast<RefIndex *>(h)->index.size();
}
const uint8_t *tref_index_data(void *h) {
        return static_cast<RefIndex *>(h)->index.data();
}
uint64_t tref_hits_size(void *h) {
        return static_cast<RefIndex *>(h)->hits.size();
}
const uint8_t *tref_hits_data(void *h) {
        return static_cast<RefIndex *>(h)->hits.data();
}
uint32_t tref_num_terms(void *h) {
        return static_cast<RefIndex *>(h)->tctx.size();
}
void tref_term(void *h, uint32_t idx, uint32_t *docs, uint32_t *off, uint32_t *len) {
        const auto &t = static_cast<RefIndex *>(h)->tctx[idx];
        *docs         = t.documents;
        *off          = t.indexChunk.offset;
        *len          = t.indexChunk.size();
}

// full decode through PostingsListIterator::next(); returns number of postings (<= cap stored)
int64_t tref_decode(void *h, uint32_t termIdx, uint32_t *docids, uint32_t *freqs, uint64_t cap) {
        auto    x = static_cast<RefIndex *>(h);
        int64_t n{0};
        if (guarded([&] {
                    std::unique_ptr<Codecs::Decoder>              dec(x->ap->new_decoder(x->tctx[termIdx]));
                    std::unique_ptr<Codecs::PostingsListIterator> it(dec->new_iterator());
                    for (auto id = it->next(); id != DocIDsEND; id = it->next()) {
                            if (uint64_t(n) < cap) {
                                    docids[n] = id;
                                    freqs[n]  = it->freq;
                            }
                            ++n;
                    }
            }))
                return -1;
        return n;
}

// every document's positions, through the reference's own PostingsListIterator::materialize_hits (google_codec.cpp:533-594,
// lucene_codec.cpp:767-856): positions[] receives freq(doc) entries per document, in list order.  Returns the number of positions.
int64_t tref_positions(void *h, uint32_t termIdx, uint32_t *positions, uint64_t cap) {
        auto    x = static_cast<RefIndex *>(h);
        int64_t n{0};
        if (guarded([&] {
                    std::unique_ptr<Codecs::Decoder>              dec(x->ap->new_decoder(x->tctx[termIdx]));
                    std::unique_ptr<Codecs::PostingsListIterator> it(dec->new_iterator());
                    DocWordsSpace                                 dws(Limits::MaxPosition);
                    std::vector<term_hit>                         hits(65536);
                    for (auto id = it->next(); id != DocIDsEND; id = it->next()) {
                            const auto freq = it->freq;
                            dws.reset();
                            it->materialize_hits(&dws, hits.data());
                            for (uint32_t i = 0; i < freq; ++i, ++n)
                                    if (uint64_t(n) < cap)
                                            positions[n] = hits[i].pos;
                    }
            }))
                return -1;
        return n;
}

// advance() probe: for each (ascending) target returns first docID >= target (DocIDsEND = UINT32_MAX when exhausted)
int tref_advance(void *h, uint32_t termIdx, const uint32_t *targets, uint32_t n, uint32_t *out) {
        auto x = static_cast<RefIndex *>(h);
        return guarded([&] {
                std::unique_ptr<Codecs::Decoder>              dec(x->ap->new_decoder(x->tctx[termIdx]));
                std::unique_ptr<Codecs::PostingsListIterator> it(dec->new_iterator());
                for (uint32_t i = 0; i < n; ++i) {
                        if (it->current() < targets[i] || it->current() == 0)
                                it->advance(targets[i]);
                        out[i] = it->current();
                }
        });
}

// reference BM25 score for (term, freq): similarity.h:209-235
double tref_bm25(void *h, uint32_t termIdx, uint32_t freq) {
        auto   x = static_cast<RefIndex *>(h);
        double r{-1};
        guarded([&] {
                Similarity::IndexSourcesCollectionBM25Scorer        cs;
                cs.reset(x->col.get());
                std::unique_ptr<Similarity::IndexSourceTermsScorer> sc(cs.new_source_scorer(x->src));
                str8_t                                              t(x->names[termIdx].data(), x->names[termIdx].size());
                std::unique_ptr<Similarity::ScorerWeight>           w(sc->new_scorer_weight(&t, 1));
                r = sc->score(1, uint16_t(freq), w.get());
        });
        return r;
}

// mode 0: ExecFlags::DocumentsOnly ; mode 1: ExecFlags::AccumulatedScoreScheme + BM25.  Returns #matches (ids/scores filled up to cap)
int64_t tref_exec2(void *h, const char *q, int mode, uint32_t parserFlags, uint32_t *ids, double *scores, uint64_t cap);
int64_t tref_exec(void *h, const char *q, int mode, uint32_t *ids, double *scores, uint64_t cap) {
        return tref_exec2(h, q, mode, 0, ids, scores, cap);
}
// the parser creates MatchSome groups ([a, b, c], ParseMatchSomeExpr = 16) with min = 1; applications raise match_some.min afterwards
static void set_match_some_min(ast_node *n, uint16_t m) {
        if (!n)
                return;
        switch (n->type) {
                case ast_node::Type::BinOp:
                        set_match_some_min(n->binop.lhs, m);
                        set_match_some_min(n->binop.rhs, m);
                        break;
                case ast_node::Type::UnaryOp:
                        set_match_some_min(n->unaryop.expr, m);
                        break;
                case ast_node::Type::ConstTrueExpr:
                        set_match_some_min(n->expr, m);
                        break;
                case ast_node::Type::MatchSome:
                        n->match_some.min = m;
                        for (size_t i = 0; i < n->match_some.size; ++i)
                                set_match_some_min(n->match_some.nodes[i], m);
                        break;
                default:
                        break;
        }
}

int64_t tref_exec3(void *h, const char *q, int mode, uint32_t parserFlags, uint32_t minMatch, uint32_t *ids, double *scores, uint64_t cap);
// parserFlags: ast_parser::Flags (queries.h:230-240), e.g. ParseConstTrueExpr = 8 enables the <expr> syntax (-> DocsSetIterators::Optional)
int64_t tref_exec2(void *h, const char *q, int mode, uint32_t parserFlags, uint32_t *ids, double *scores, uint64_t cap) {
        return tref_exec3(h, q, mode, parserFlags, 0, ids, scores, cap);
}
// minMatch != 0: match_some.min of every MatchSome group of the parsed query
int64_t tref_exec3(void *h, const char *q, int mode, uint32_t parserFlags, uint32_t minMatch, uint32_t *ids, double *scores, uint64_t cap) {
        auto    x = static_cast<RefIndex *>(h);
        int64_t n{-1};
        guarded([&] {
                query       qq(str32_t(q, strlen(q)), default_token_parser_impl, parserFlags);
                if (minMatch)
                        set_match_some_min(qq.root, uint16_t(minMatch));
                CollectSink sink;
                sink.cap  = cap;
                auto reg  = masked_documents_registry::make(nullptr, 0);
                if (mode == 0) {
                        exec_query(qq, x->src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::DocumentsOnly));
                } else {
                        Similarity::IndexSourcesCollectionBM25Scorer        cs;
                        cs.reset(x->col.get());
                        std::unique_ptr<Similarity::IndexSourceTermsScorer> sc(cs.new_source_scorer(x->src));
                        exec_query(qq, x->src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::AccumulatedScoreScheme), sc.get());
                }
                memcpy(ids, sink.ids.data(), sink.ids.size() * sizeof(uint32_t));
                if (mode == 1 && scores)
                        memcpy(scores, sink.scores.data(), sink.scores.size() * sizeof(double));
                n = int64_t(sink.n);
        });
        return n;
}

// exec_query with a masked_documents_registry built from `masked` docIDs through the reference's own pack_updates / unpack_updates
// (docidupdates.cpp:8-118) — the per-match maskedDocumentsRegistry->test(id) of the exec Handlers (exec.cpp:1108-1116)
int64_t tref_exec_masked(void *h, const char *q, int mode, const uint32_t *masked, uint32_t nmasked, uint32_t *ids, double *scores, uint64_t cap) {
        auto    x = static_cast<RefIndex *>(h);
        int64_t n{-1};
        guarded([&] {
                query                qq(str32_t(q, strlen(q)));
                CollectSink          sink;
                std::vector<docid_t> v(masked, masked + nmasked);
                IOBuffer             packed;
                sink.cap = cap;
                pack_updates(v, &packed);
                auto ud  = unpack_updates({reinterpret_cast<const uint8_t *>(packed.data()), uint32_t(packed.size())});
                auto reg = masked_documents_registry::make(&ud, 1);
                if (mode == 0) {
                        exec_query(qq, x->src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::DocumentsOnly));
                } else {
                        Similarity::IndexSourcesCollectionBM25Scorer        cs;
                        cs.reset(x->col.get());
                        std::unique_ptr<Similarity::IndexSourceTermsScorer> sc(cs.new_source_scorer(x->src));
                        exec_query(qq, x->src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::AccumulatedScoreScheme), sc.get());
                }
                memcpy(ids, sink.ids.data(), sink.ids.size() * sizeof(uint32_t));
                if (mode == 1 && scores)
                        memcpy(scores, sink.scores.data(), sink.scores.size() * sizeof(double));
                n = int64_t(sink.n);
        });
        return n;
}

// ---- segments: written by the reference's own SegmentIndexSession (indexer.cpp) and opened by its SegmentIndexSource
// doc-major input is assembled from term-major lists; documents with id < replace_below are replace()d, the others insert()ed; `dir` must end in a numeric generation (segment_index_source.cpp:18-21)
int tref_segment_write(int codec, const char *dir, uint32_t nterms, const char *const *names, const uint32_t *counts, const uint32_t *docids, const uint32_t *freqs,
                       const uint32_t *erased, uint32_t nerased, uint32_t replace_below) {
        return guarded([&] {
                struct Hit {
                        uint32_t doc, term, freq;
                };
                std::vector<Hit> hits;
                size_t           at{0};
                for (uint32_t t = 0; t < nterms; ++t)
                        for (uint32_t i = 0; i < counts[t]; ++i, ++at)
                                hits.push_back({docids[at], t, freqs[at]});
                std::sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.doc != b.doc ? a.doc < b.doc : a.term < b.term; });
                SegmentIndexSession sess;
                for (size_t i = 0; i < hits.size();) {
                        const auto doc   = hits[i].doc;
                        auto       proxy = sess.begin(doc);
                        tokenpos_t pos{1};
                        for (; i < hits.size() && hits[i].doc == doc; ++i) {
                                const str8_t term(names[hits[i].term], uint8_t(strlen(names[hits[i].term])));
                                if (hits[i].freq == 0)
                                        proxy.insert(term, 0); // a posting without positional hits (freq 0)
                                for (uint32_t k = 0; k < hits[i].freq; ++k)
                                        proxy.insert(term, pos++);
                        }
                        if (doc < replace_below)
                                sess.replace(proxy); // a document an older segment already holds: recorded in updated_documents.ids
                        else
                                sess.insert(proxy);
                }
                for (uint32_t i = 0; i < nerased; ++i)
                        sess.erase(erased[i]);
                if (codec == 0) {
                        Codecs::Google::IndexSession cs(dir);
                        sess.commit(&cs);
                } else {
                        Codecs::Lucene::IndexSession cs(dir);
                        sess.commit(&cs);
                }
        });
}

void *tref_segment_open(const char *dir) {
        auto x = new RefIndex();
        if (guarded([&] {
                    auto seg = new SegmentIndexSource(dir);
                    x->src   = seg;
                    x->codec = -1;
                    x->col.reset(new IndexSourcesCollection());
                    x->col->insert(seg);
                    seg->Release();
                    x->col->commit();
            })) {
                delete x;
                return nullptr;
        }
        return x;
}

// several segments as one IndexSourcesCollection (index_source.cpp:3-30: newest generation first; source i is scanned with the
// updated_documents of every NEWER source as its masked_documents_registry) — what Trinity applications loop over per query
void *tref_collection_open(const char *const *dirs, uint32_t n) {
        auto x = new RefIndex();
        if (guarded([&] {
                    x->codec = -1;
                    x->col.reset(new IndexSourcesCollection());
                    for (uint32_t i = 0; i < n; ++i) {
                            auto seg = new SegmentIndexSource(dirs[i]);
                            x->col->insert(seg);
                            seg->Release();
                    }
                    x->col->commit();
                    x->src = x->col->sources.front();
            })) {
                delete x;
                return nullptr;
        }
        return x;
}

// exec_query over every source of the collection; results are concatenated in collection order, seg_counts[i] = #matches of source i
int64_t tref_collection_exec(void *h, const char *q, int mode, uint32_t *ids, double *scores, uint64_t cap, uint64_t *seg_counts) {
        auto    x = static_cast<RefIndex *>(h);
        int64_t n{-1};
        guarded([&] {
                query                                        qq(str32_t(q, strlen(q)));
                Similarity::IndexSourcesCollectionBM25Scorer cs;
                uint64_t                                     at{0};
                cs.reset(x->col.get());
                for (size_t i = 0; i < x->col->sources.size(); ++i) {
                        auto        src = x->col->sources[i];
                        auto        reg = x->col->scanner_registry_for(uint16_t(i));
                        CollectSink sink;
                        sink.cap = cap - at;
                        if (mode == 0) {
                                exec_query(qq, src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::DocumentsOnly));
                        } else {
                                std::unique_ptr<Similarity::IndexSourceTermsScorer> sc(cs.new_source_scorer(src));
                                exec_query(qq, src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::AccumulatedScoreScheme), sc.get());
                        }
                        memcpy(ids + at, sink.ids.data(), sink.ids.size() * sizeof(uint32_t));
                        if (mode == 1 && scores)
                                memcpy(scores + at, sink.scores.data(), sink.scores.size() * sizeof(double));
                        if (seg_counts)
                                seg_counts[i] = sink.n;
                        at += sink.ids.size();
                }
                n = int64_t(at);
        });
        return n;
}

int tref_resolve(void *h, const char *term, uint32_t *docs, uint32_t *off, uint32_t *len) {
        auto x = static_cast<RefIndex *>(h);
        return guarded([&] {
                const auto t = x->src->resolve_term_ctx(str8_t(term, uint8_t(strlen(term))));
                *docs        = t.documents;
                *off         = t.indexChunk.offset;
                *len         = t.indexChunk.size();
        });
}

int tref_field_stats(void *h, uint64_t *sumTermHits, uint32_t *totalTerms, uint64_t *sumTermsDocs, uint32_t *docsCnt) {
        auto x = static_cast<RefIndex *>(h);
        return guarded([&] {
                const auto fs = x->src->default_field_stats();
                *sumTermHits  = fs.sumTermHits;
                *totalTerms   = fs.totalTerms;
                *sumTermsDocs = fs.sumTermsDocs;
                *docsCnt      = fs.docsCnt;
        });
}

// CPU baseline: run nq queries over `threads` host threads (one query per thread at a time; exec_query is re-entrant, exec.cpp:12).
// mode 0: collect every matched docID (DocumentsOnly); mode 1: BM25 top-k heap in consider(id, score).
// Outputs per query: match count, sum of matched ids (mode 0) , top-k (mode 1: ids/scores, k per query, padded with 0).
// Returns elapsed wall seconds for the whole batch (steady_clock), < 0 on error.
double tref_exec_batch(void *h, const char *const *qs, uint32_t nq, int mode, uint32_t k, int threads, uint64_t *matchCounts, uint64_t *idSums,
                       uint32_t *topkIds, double *topkScores) {
        auto                  x = static_cast<RefIndex *>(h);
        std::atomic<uint32_t> next{0};
        std::atomic<int>      failed{0};
        std::string           err;
        std::mutex            errLock;
        const auto            t0 = std::chrono::steady_clock::now();
        auto                  worker = [&] {
                try {
                        IdsSink                                             ids;
                        Similarity::IndexSourcesCollectionBM25Scorer        cs;
                        std::unique_ptr<Similarity::IndexSourceTermsScorer> sc;
                        if (mode == 1) {
                                cs.reset(x->col.get());
                                sc.reset(cs.new_source_scorer(x->src));
                        }
                        auto reg = masked_documents_registry::make(nullptr, 0);
                        for (;;) {
                                const auto i = next.fetch_add(1);
                                if (i >= nq)
                                        break;
                                query qq(str32_t(qs[i], strlen(qs[i])));
                                if (mode == 0) {
                                        ids.ids.clear();
                                        ids.sum = 0;
                                        exec_query(qq, x->src, reg.get(), &ids, nullptr, uint32_t(ExecFlags::DocumentsOnly));
                                        if (matchCounts)
                                                matchCounts[i] = ids.ids.size();
                                        if (idSums)
                                                idSums[i] = ids.sum;
                                } else {
                                        TopKSink sink;
                                        sink.k = k;
                                        exec_query(qq, x->src, reg.get(), &sink, nullptr, uint32_t(ExecFlags::AccumulatedScoreScheme), sc.get());
                                        if (matchCounts)
                                                matchCounts[i] = sink.n;
                                        size_t cnt = sink.pq.size();
                                        for (uint32_t j = 0; j < k; ++j) {
                                                if (topkIds)
                                                        topkIds[size_t(i) * k + j] = 0;
                                                if (topkScores)
                                                        topkScores[size_t(i) * k + j] = 0;
                                        }
                                        while (cnt) {
                                                --cnt;
                                                if (topkIds)
                                                        topkIds[size_t(i) * k + cnt] = sink.pq.top().id;
                                                if (topkScores)
                                                        topkScores[size_t(i) * k + cnt] = sink.pq.top().s;
                                                sink.pq.pop();
                                        }
                                }
                        }
                } catch (const std::exception &e) {
                        std::lock_guard<std::mutex> g(errLock);
                        err = e.what();
                        failed.store(1);
                } catch (...) {
                        failed.store(1);
                }
        };
        std::vector<std::thread> ths;
        if (threads < 1)
                threads = 1;
        for (int i = 1; i < threads; ++i)
                ths.emplace_back(worker);
        worker();
        for (auto &t : ths)
                t.join();
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (failed.load()) {
                g_err = err.empty() ? "exec_query failed" : err;
                return -1.0;
        }
        return el;
}
}


// ------------------------------------------------------------------------------------------------------------------------------
// The BASELINE.md synthetic Zipfian index authored by the REFERENCE's own Encoders, so that bench.py's --impl reference
// arm never maps the product library: V terms, df_r = max(min_df, floor(0.5 * N / r)), geometric docID gaps from splitmix64(seed ^ r),
// freq = 1 + min(7, Geom(1/2)), positions cumulative 2..17 — the same workload generator as trn_synth_build (restated here; the two are
// held byte-equal by tests/test_codecs_cpu.py).  Terms are encoded in parallel, one reference IndexSession per term; the one piece of
// encoder state that survives end_term() — the Google skiplist countdown (google_codec.h:57) — is reproduced through the public API by
// encoding a throw-away term with (blocks committed so far mod 8) blocks first.
namespace {
        inline uint64_t sm64(uint64_t &s) {
                uint64_t z = (s += 0x9E3779B97F4A7C15ull);
                z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                return z ^ (z >> 31);
        }
        inline uint32_t synth_df(uint32_t ndocs, uint32_t rank, uint32_t min_df) {
                const uint64_t z = uint64_t(ndocs) / (2ull * rank);
                return uint32_t(std::min<uint64_t>(ndocs, std::max<uint64_t>(min_df, z)));
        }
        template <class F>
        void synth_term(uint32_t ndocs, uint32_t rank, uint32_t min_df, uint64_t seed, F &&f) {
                const uint32_t df = synth_df(ndocs, rank, min_df);
                uint64_t       s  = seed ^ uint64_t(rank);
                uint64_t       s2 = (seed ^ uint64_t(rank)) * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;
                const double   p  = double(df) / double(ndocs);
                const double   il = p < 1.0 ? 1.0 / std::log1p(-p) : 0.0;
                uint64_t       doc{0};
                for (uint32_t i = 0; i < df; ++i) {
                        const uint64_t x = sm64(s);
                        uint64_t       gap{1};
                        if (p < 1.0) {
                                const double u = double((x >> 11) + 1) * (1.0 / 9007199254740992.0);
                                const double g = std::floor(std::log(u) * il);
                                gap            = 1 + uint64_t(std::min(g, 4.0e9));
                        }
                        const uint64_t maxdoc = uint64_t(ndocs) - (df - 1 - i);
                        doc                   = std::min(doc + gap, maxdoc);
                        const uint32_t geo    = uint32_t(__builtin_ctzll((x & 0x7ffull) | 0x800ull));
                        f(uint32_t(doc), 1 + std::min<uint32_t>(7, geo), sm64(s2));
                }
        }
} // namespace

extern "C" void *tref_synth_build(int codec, uint32_t ndocs, uint32_t nterms, uint32_t min_df, uint64_t seed, int with_hits, int threads) {
        struct Part {
                std::vector<uint8_t> index, hits;
                uint32_t             documents{0};
                uint64_t             hitsCnt{0};
        };
        std::vector<Part>     parts(nterms);
        std::vector<uint32_t> warm(nterms, 0); // Google: blocks the throw-away term must commit first
        if (codec == 0) {
                uint64_t blocks{0};
                for (uint32_t r = 1; r <= nterms; ++r) {
                        warm[r - 1] = uint32_t(blocks % Codecs::Google::SKIPLIST_STEP);
                        blocks += (synth_df(ndocs, r, min_df) + Codecs::Google::N - 1) / Codecs::Google::N;
                }
        }
        std::atomic<uint32_t> next{0};
        std::atomic<int>      failed{0};
        auto                  worker = [&] {
                try {
                        for (;;) {
                                const uint32_t i = next.fetch_add(1);
                                if (i >= nterms || failed.load())
                                        break;
                                std::unique_ptr<Codecs::IndexSession> sess;
                                if (codec == 0)
                                        sess.reset(new Codecs::Google::IndexSession("/tmp"));
                                else
                                        sess.reset(new Codecs::Lucene::IndexSession("/tmp"));
                                sess->begin();
                                std::unique_ptr<Codecs::Encoder> enc(sess->new_encoder());
                                term_index_ctx                   t;
                                if (warm[i]) { // full blocks except the last, which end_term() commits with a single document
                                        enc->begin_term();
                                        const uint32_t ndummy = uint32_t(Codecs::Google::N) * (warm[i] - 1) + 1;
                                        for (uint32_t d = 1; d <= ndummy; ++d) {
                                                enc->begin_document(d);
                                                enc->new_hit(1, {});
                                                enc->end_document();
                                        }
                                        enc->end_term(&t);
                                }
                                auto &P = parts[i];
                                enc->begin_term();
                                synth_term(ndocs, i + 1, min_df, seed, [&](uint32_t doc, uint32_t freq, uint64_t y) {
                                        enc->begin_document(doc);
                                        uint32_t pos{0};
                                        for (uint32_t h = 0; h < freq; ++h) {
                                                pos = with_hits ? pos + 2u + uint32_t((y >> (4u * h)) & 15u) : h + 1;
                                                enc->new_hit(pos, {});
                                        }
                                        P.hitsCnt += freq;
                                        enc->end_document();
                                });
                                enc->end_term(&t);
                                const auto *ib = reinterpret_cast<const uint8_t *>(sess->indexOut.data());
                                P.index.assign(ib + t.indexChunk.offset, ib + t.indexChunk.offset + t.indexChunk.size());
                                P.documents = t.documents;
                                if (codec == 1) {
                                        auto ls = static_cast<Codecs::Lucene::IndexSession *>(sess.get());
                                        P.hits.assign(reinterpret_cast<const uint8_t *>(ls->positionsOut.data()),
                                                      reinterpret_cast<const uint8_t *>(ls->positionsOut.data()) + ls->positionsOut.size());
                                }
                        }
                } catch (...) {
                        failed.store(1);
                }
        };
        std::vector<std::thread> ths;
        for (int i = 1; i < std::max(1, threads); ++i)
                ths.emplace_back(worker);
        worker();
        for (auto &t : ths)
                t.join();
        if (failed.load()) {
                g_err = "tref_synth_build: a reference encoder failed";
                return nullptr;
        }
        uint64_t ib{0}, hb{0};
        for (auto &p : parts) {
                ib += p.index.size();
                hb += p.hits.size();
        }
        if (ib >= (1ull << 32) || hb >= (1ull << 32)) {
                g_err = "tref_synth_build: index larger than range32_t";
                return nullptr;
        }
        auto x   = new RefIndex();
        x->codec = codec;
        x->index.resize(ib);
        x->hits.resize(hb);
        uint64_t io{0}, ho{0};
        char     name[16];
        for (uint32_t i = 0; i < nterms; ++i) {
                auto &p = parts[i];
                std::memcpy(x->index.data() + io, p.index.data(), p.index.size());
                if (codec == 1) { // the chunk header's first u32 is the term's absolute offset into hits.data (lucene_codec.cpp:178)
                        const uint32_t v = uint32_t(ho);
                        std::memcpy(x->index.data() + io, &v, 4);
                        if (!p.hits.empty())
                                std::memcpy(x->hits.data() + ho, p.hits.data(), p.hits.size());
                }
                snprintf(name, sizeof(name), "t%04u", i + 1);
                x->names.emplace_back(name);
                x->tctx.emplace_back(p.documents, range32_t{uint32_t(io), uint32_t(p.index.size())});
                x->sumHits += p.hitsCnt;
                io += p.index.size();
                ho += p.hits.size();
                std::vector<uint8_t>().swap(p.index);
                std::vector<uint8_t>().swap(p.hits);
        }
        if (guarded([&] { x->open(ndocs); })) {
                delete x;
                return nullptr;
        }
        return x;
}


// ------------------------------------------------------------------------------------------------------------------------------
// Only in oracle/_ref/libtrinity_ref_gpu.so (built with -DTRINITY_B200_GPU_SPAN): attach a device-resident twin (integration/gpu_exec.h,
// the reference-side binding of libtrinity_b200.so) to the index source, so that the library's exec_query() — the reference's own, with
// its one span-building call site going through b200_gpu_span() — executes on the GPU and replays into the same Handlers / consider().
#ifdef TRINITY_B200_GPU_SPAN
#include "gpu_exec.h"
namespace {
        std::unordered_map<void *, std::unique_ptr<Trinity::GpuAccessProxy>> g_gaps;
}
extern "C" int tref_gpu_attach2(void *h, int device, uint64_t maxDocID, int withHits);
extern "C" int tref_gpu_attach(void *h, int device, uint64_t maxDocID) {
        return tref_gpu_attach2(h, device, maxDocID, 1);
}
extern "C" int tref_gpu_attach2(void *h, int device, uint64_t maxDocID, int withHits) { // withHits = 0: a LUCENE source whose hits.data stays on the host
        auto x = static_cast<RefIndex *>(h);
        return guarded([&] {
                std::vector<std::pair<std::string, term_index_ctx>> terms;
                for (size_t i = 0; i < x->names.size(); ++i)
                        terms.emplace_back(x->names[i], x->tctx[i]);
                const bool hits = withHits && !x->hits.empty();
                auto       gap  = std::make_unique<Trinity::GpuAccessProxy>(device, x->ap.get(), x->index.size(), terms, isrc_docid_t(maxDocID), hits ? x->hits.data() : nullptr,
                                                                     hits ? x->hits.size() : 0);
                Trinity::gpu_proxy_register(x->src, gap.get());
                g_gaps[h] = std::move(gap);
        });
}
extern "C" void tref_gpu_detach(void *h) {
        auto x = static_cast<RefIndex *>(h);
        Trinity::gpu_proxy_register(x->src, nullptr);
        g_gaps.erase(h);
}
extern "C" uint64_t tref_gpu_spans_executed(void *h) {
        const auto it = g_gaps.find(h);
        return it == g_gaps.end() ? 0 : it->second->spansExecuted;
}
#endif
