// See gpu_exec.h.  Reference-side binding: everything here speaks the reference's types on one side and the C ABI on the other.
#include "gpu_exec.h"
#include "lucene_codec.h"
#include <mutex>
#include <stdexcept>

namespace Trinity {

GpuAccessProxy::GpuAccessProxy(int device, Codecs::AccessProxy *ap, size_t indexSize, const std::vector<std::pair<std::string, term_index_ctx>> &terms, isrc_docid_t maxDocID,
                               const uint8_t *hitsData, size_t hitsSize) {
        if (trn_create(device, &ctx) != TRN_OK) {
                const std::string m = ctx ? trn_last_error(ctx) : "trn_create failed";
                if (ctx)
                        trn_destroy(ctx);
                throw Switch::system_error(m.c_str()); // errors -> the reference's exception types (exec.h:46-47)
        }
        std::vector<trn_term> t;
        t.reserve(terms.size());
        for (const auto &[term, tctx] : terms) {
                idOf.emplace(term, uint32_t(t.size()));
                t.push_back({tctx.documents, tctx.indexChunk.offset, tctx.indexChunk.size()});
        }
        const auto ci    = ap->codec_identifier();
        codec            = (ci.size() == 6 && !memcmp(ci.data(), "GOOGLE", 6)) ? TRN_CODEC_GOOGLE : TRN_CODEC_LUCENE;
        if (trn_upload_index(ctx, codec, ap->indexPtr, indexSize, t.data(), uint32_t(t.size()), maxDocID) != TRN_OK) {
                const std::string m = trn_last_error(ctx);
                trn_destroy(ctx);
                throw Switch::data_error(m.c_str());
        }
        // LUCENE keeps positions in hits.data (Lucene::AccessProxy::hitsDataPtr, lucene_codec.h:204-218): with them on the device the span
        // takes phrase plans too; without them those stay with the reference's own span
        if (codec == TRN_CODEC_LUCENE) {
                if (!hitsData) {
                        const auto lap = static_cast<Codecs::Lucene::AccessProxy *>(ap);
                        hitsData       = lap->hitsDataPtr;
                        hitsSize       = lap->hitsDataSize; // (0 when the proxy was handed a pointer: pass the size to this constructor then)
                }
                if (hitsData && hitsSize) {
                        if (trn_upload_hits(ctx, ap->indexPtr, indexSize, hitsData, hitsSize) != TRN_OK) {
                                const std::string m = trn_last_error(ctx);
                                trn_destroy(ctx);
                                throw Switch::data_error(m.c_str());
                        }
                        havePositions = true;
                }
        } else
                havePositions = true; // GOOGLE: inline
}

GpuAccessProxy::~GpuAccessProxy() {
        trn_destroy(ctx);
}

namespace {
        std::mutex                                         g_regLock;
        std::unordered_map<IndexSource *, GpuAccessProxy *> g_registry;
} // namespace

void gpu_proxy_register(IndexSource *src, GpuAccessProxy *gap) {
        std::lock_guard<std::mutex> g(g_regLock);
        if (gap)
                g_registry[src] = gap;
        else
                g_registry.erase(src);
}

GpuAccessProxy *gpu_proxy_for(IndexSource *src) {
        std::lock_guard<std::mutex> g(g_regLock);
        const auto                  it = g_registry.find(src);
        return it == g_registry.end() ? nullptr : it->second;
}

namespace {
        // exec_node tree (compilation_ctx.h:8-165) -> plan tree -> trn_qnode[] (children contiguous, after their parent).
        // Mirrors queryexec_ctx::build_iterator case by case, including the way it pulls the operands of a nested conjunction /
        // disjunction into the parent (exec.cpp:328-400).
        struct PlanBuilder {
                struct PNode {
                        uint8_t          kind{TRN_NODE_TERM};
                        uint32_t         term{0xffffffffu}; // TERM: trn term id (0xffffffff: the source does not hold the term); SOME: min
                        double           weight{0};
                        std::vector<int> kids;
                };
                queryexec_ctx &                     rctx;
                GpuAccessProxy &                    gap;
                Similarity::IndexSourceTermsScorer *scorer;
                std::vector<PNode>                  n;
                bool                                unsupported{false};

                int term_node(const exec_term_id_t termID) {
                        const auto &info = rctx.tctxMap[termID];
                        PNode       x;
                        const auto  it = gap.idOf.find(std::string(info.second.data(), info.second.size()));
                        x.term         = it == gap.idOf.end() ? 0xffffffffu : it->second;
                        if (scorer) { // == the PLI wrapper of docset_iterators_scorers.cpp: one ScorerWeight per term instance
                                auto                                      token = info.second;
                                std::unique_ptr<Similarity::ScorerWeight> w(scorer->new_scorer_weight(&token, 1));
                                x.weight = static_cast<Similarity::IndexSourcesCollectionBM25Scorer::Scorer::ScorerWeight *>(w.get())->idf;
                        }
                        n.push_back(x);
                        return int(n.size()) - 1;
                }
                // the engine checks positions on the device: GOOGLE has them inline, a LUCENE source needs its hits.data uploaded (constructor);
                // without it such plans stay with the reference's own span
                int phrase_node(const compilation_ctx::phrase *p) {
                        if (!gap.havePositions) {
                                unsupported = true;
                                return term_node(p->termIDs[0]);
                        }
                        std::vector<int> kids;
                        for (uint8_t i = 0; i < p->size; ++i)
                                kids.push_back(term_node(p->termIDs[i]));
                        return group(TRN_NODE_PHRASE, std::move(kids));
                }
                int group(uint8_t kind, std::vector<int> kids, uint32_t min = 0) {
                        PNode x;
                        x.kind = kind;
                        x.term = min;
                        x.kids = std::move(kids);
                        n.push_back(std::move(x));
                        return int(n.size()) - 1;
                }
                // operands of a conjunction (and = true) / disjunction rooted at e, nested same-kind operators pulled up
                void collect(const exec_node e, const bool isAnd, std::vector<int> &out) {
                        const auto allT = isAnd ? ENT::matchallterms : ENT::matchanyterms;
                        const auto allN = isAnd ? ENT::matchallnodes : ENT::matchanynodes;
                        const auto bin  = isAnd ? ENT::logicaland : ENT::logicalor;
                        if (e.fp == allT) {
                                const auto run = static_cast<const compilation_ctx::termsrun *>(e.ptr);
                                for (uint16_t i = 0; i < run->size; ++i)
                                        out.push_back(term_node(run->terms[i]));
                        } else if (e.fp == allN) {
                                const auto g = static_cast<const compilation_ctx::nodes_group *>(e.ptr);
                                for (uint16_t i = 0; i < g->size; ++i)
                                        collect(g->nodes[i], isAnd, out);
                        } else if (e.fp == bin) {
                                const auto b = static_cast<const compilation_ctx::binop_ctx *>(e.ptr);
                                if (isAnd && (b->lhs.fp == ENT::consttrueexpr || b->rhs.fp == ENT::consttrueexpr)) {
                                        out.push_back(emit(e)); // Optional(main, opt): an operand of its own
                                        return;
                                }
                                collect(b->lhs, isAnd, out);
                                collect(b->rhs, isAnd, out);
                        } else if (e.fp == ENT::unaryand || e.fp == ENT::consttrueexpr) {
                                collect(static_cast<const compilation_ctx::unaryop_ctx *>(e.ptr)->expr, isAnd, out);
                        } else
                                out.push_back(emit(e));
                }
                int emit(const exec_node e) {
                        switch (e.fp) {
                                case ENT::matchterm:
                                        return term_node(e.u16);
                                case ENT::matchallterms:
                                case ENT::matchallnodes: {
                                        std::vector<int> kids;
                                        collect(e, true, kids);
                                        return kids.size() == 1 ? kids[0] : group(TRN_NODE_AND, std::move(kids));
                                }
                                case ENT::matchanyterms:
                                case ENT::matchanynodes:
                                case ENT::logicalor: {
                                        std::vector<int> kids;
                                        collect(e, false, kids);
                                        return kids.size() == 1 ? kids[0] : group(TRN_NODE_OR, std::move(kids));
                                }
                                case ENT::logicaland: {
                                        const auto b = static_cast<const compilation_ctx::binop_ctx *>(e.ptr);
                                        if (b->lhs.fp == ENT::consttrueexpr || b->rhs.fp == ENT::consttrueexpr) { // -> Optional(main, opt), exec.cpp:370-377
                                                const bool l   = b->lhs.fp == ENT::consttrueexpr;
                                                const auto opt = static_cast<const compilation_ctx::unaryop_ctx *>((l ? b->lhs : b->rhs).ptr)->expr;
                                                const int  m   = emit(l ? b->rhs : b->lhs);
                                                const int  o   = emit(opt);
                                                return group(TRN_NODE_OPTIONAL, {m, o});
                                        }
                                        std::vector<int> kids;
                                        collect(e, true, kids);
                                        return kids.size() == 1 ? kids[0] : group(TRN_NODE_AND, std::move(kids));
                                }
                                case ENT::logicalnot: { // -> Filter(req = lhs, excl = rhs)
                                        const auto b = static_cast<const compilation_ctx::binop_ctx *>(e.ptr);
                                        const int  r = emit(b->lhs);
                                        const int  x = emit(b->rhs);
                                        return group(TRN_NODE_NOT, {r, x});
                                }
                                case ENT::unaryand:
                                case ENT::consttrueexpr:
                                        return emit(static_cast<const compilation_ctx::unaryop_ctx *>(e.ptr)->expr);
                                case ENT::matchsome: { // -> DisjunctionSome(nodes, min)
                                        const auto       g = static_cast<const compilation_ctx::partial_match_ctx *>(e.ptr);
                                        std::vector<int> kids;
                                        for (uint16_t i = 0; i < g->size; ++i)
                                                kids.push_back(emit(g->nodes[i]));
                                        return group(TRN_NODE_SOME, std::move(kids), g->min);
                                }
                                case ENT::matchphrase: // -> Phrase(PLIs in phrase order), exec.cpp:284-297
                                        return phrase_node(static_cast<const compilation_ctx::phrase *>(e.ptr));
                                case ENT::matchanyphrases:   // -> Disjunction of Phrases, exec.cpp:298-312
                                case ENT::matchallphrases: { // -> Conjuction of Phrases, exec.cpp:313-327
                                        const auto       run = static_cast<const compilation_ctx::phrasesrun *>(e.ptr);
                                        std::vector<int> kids;
                                        for (uint16_t i = 0; i < run->size; ++i)
                                                kids.push_back(phrase_node(run->phrases[i]));
                                        return kids.size() == 1 ? kids[0] : group(e.fp == ENT::matchallphrases ? TRN_NODE_AND : TRN_NODE_OR, std::move(kids));
                                }
                                default: // anything else (constfalse survivors, ...) stays with the reference's own span
                                        unsupported = true;
                                        return term_node(0);
                        }
                }
                // breadth-first serialisation: children contiguous and behind their parent
                std::vector<trn_qnode> serialise(int root) const {
                        std::vector<trn_qnode> out(1);
                        std::vector<int>       order{root};
                        for (size_t qi = 0; qi < order.size(); ++qi) {
                                const PNode &A = n[order[qi]];
                                trn_qnode    q;
                                memset(&q, 0, sizeof(q));
                                q.kind = A.kind;
                                if (A.kind == TRN_NODE_TERM) {
                                        q.term   = A.term;
                                        q.weight = A.weight;
                                } else {
                                        q.term        = A.term;
                                        q.nchildren   = uint8_t(A.kids.size());
                                        q.first_child = uint16_t(out.size());
                                        for (int k : A.kids) {
                                                order.push_back(k);
                                                out.emplace_back();
                                        }
                                }
                                out[qi] = q;
                        }
                        return out;
                }
        };

        // The batch operator: a DocsSetSpan whose process() is ONE trn_exec_batch call followed by the MatchesProxy replay.
        struct GpuDocsSetSpan final : public DocsSetSpan {
                GpuAccessProxy &       gap;
                std::vector<trn_qnode> plan;
                const bool             scored;

                GpuDocsSetSpan(GpuAccessProxy &g, std::vector<trn_qnode> p, bool s)
                    : gap{g}, plan{std::move(p)}, scored{s} {
                }

                uint64_t cost() override final { // docset_spans.h:89; nobody wraps this span, the value only has to be an upper bound
                        return DocIDsEND;
                }

                isrc_docid_t process(MatchesProxy *mp, const isrc_docid_t min, const isrc_docid_t max) override final {
                        trn_query  q{plan.data(), uint32_t(plan.size()), 0};
                        trn_result r;
                        // DocumentsOnly: the compact result form (a docID tile's matches as bitmap / bucketed 8-bit offsets / 16-bit offsets / docIDs) — what a batch
                        // of queries would use to keep the host link out of the way; replayed below by trn_result_for_each
                        if (trn_exec_batch(gap.ctx, &q, 1, scored ? TRN_MODE_SCORED_ALL : TRN_MODE_DOCS_COMPACT, 0, &r) != TRN_OK)
                                throw Switch::system_error(trn_last_error(gap.ctx));
                        ++gap.spansExecuted;
                        relevant_document rd; // docset_iterators.h:456-497: carries (id, score_) to the Handler
                        if (!scored) {
                                struct Replay {
                                        MatchesProxy *     mp;
                                        relevant_document *rd;
                                        isrc_docid_t       min, max, stoppedAt;
                                } st{mp, &rd, min, max, DocIDsEND};
                                // ascending docID, exactly once per match; an exception thrown by consider() (aborted_search_exception) unwinds
                                // through the replay like it unwinds through the reference's own span
                                const int rc = trn_result_for_each(&r, 0, [](void *ctx, uint32_t id) -> int {
                                        auto &S = *static_cast<Replay *>(ctx);
                                        if (id < S.min)
                                                return 0;
                                        if (id >= S.max) {
                                                S.stoppedAt = id;
                                                return 1;
                                        }
                                        S.rd->set_document(id);
                                        S.mp->process(S.rd); // -> Handler::process -> maskedDocs / consider(id) (exec.cpp:1095-1345)
                                        return 0;
                                }, &st);
                                if (rc != TRN_OK)
                                        throw Switch::data_error("malformed compact result");
                                return st.stoppedAt;
                        }
                        for (uint64_t i = r.offsets[0]; i < r.offsets[1]; ++i) { // ascending docID, exactly once per match
                                const auto id = r.docids[i];
                                if (id < min)
                                        continue;
                                if (id >= max)
                                        return id;
                                rd.set_document(id);
                                rd.score_ = r.scores[i];
                                mp->process(&rd); // -> Handler::process -> maskedDocs / consider(id, score) (exec.cpp:1095-1345)
                        }
                        return DocIDsEND;
                }
        };
} // namespace

std::unique_ptr<DocsSetSpan> b200_gpu_span(queryexec_ctx &rctx, const exec_node root, const uint32_t execFlags, IndexSource *idxsrc, Similarity::IndexSourceTermsScorer *scorer) {
        const bool documentsOnly = execFlags & uint32_t(ExecFlags::DocumentsOnly), accum = execFlags & uint32_t(ExecFlags::AccumulatedScoreScheme);
        if (!documentsOnly && !accum)
                return nullptr; // the default (rich matched_document) mode stays on the CPU
        auto gap = gpu_proxy_for(idxsrc);
        if (!gap)
                return nullptr;
        PlanBuilder pb{rctx, *gap, accum ? scorer : nullptr, {}, false};
        const int   r = pb.emit(root);
        if (pb.unsupported)
                return nullptr;
        return std::make_unique<GpuDocsSetSpan>(*gap, pb.serialise(r), accum);
}

} // namespace Trinity
