// Reference-side binding of libtrinity_b200.so: the code a Trinity maintainer adds to the Trinity tree.
// It compiles against the reference's own headers (exec.h, queryexec_ctx.h, ...) and the C ABI (include/trinity_b200.h) and is
// built — together with the one-call-site patch of exec_query() — by oracle/build_ref.sh into oracle/_ref/libtrinity_ref_gpu.so, so that
// the "drops in under exec_query()" claim is executed by tests/test_gpu_boundary.py, not just written down.
//
//   GpuAccessProxy   == Codecs::AccessProxy for a device-resident index (codecs.h:290-317): uploads indexPtr once
//   PlanBuilder      == queryexec_ctx::build_iterator (exec.cpp:253-449): exec_node tree -> trn_qnode[] (same flattening rules)
//   GpuDocsSetSpan   == a DocsSetSpan (docset_spans.h:80-89) whose process() is ONE trn_exec_batch call + the MatchesProxy replay
//   b200_gpu_span()  == the hook exec_query() calls instead of build_iterator + build_span (exec.cpp:1083-1086)
#pragma once
#include "docset_spans.h"
#include "exec.h"
#include "queryexec_ctx.h"
#include "similarity.h"
#include "../include/trinity_b200.h"
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace Trinity {

// One per (IndexSource, GPU).  Owns the device copy of AccessProxy::indexPtr.
struct GpuAccessProxy final {
        trn_ctx *                                 ctx{nullptr};
        std::unordered_map<std::string, uint32_t> idOf; // term -> trn term id
        uint64_t                                  spansExecuted{0};
        int                                       codec{0}; // TRN_CODEC_*
        bool                                      havePositions{false}; // GOOGLE: inline hits; LUCENE: hits.data uploaded

        // `terms`: every (term, term_index_ctx) of the source, e.g. from SegmentTerms iteration (terms.h:27-37)
        // LUCENE: hits.data is taken from the Lucene::AccessProxy (hitsDataPtr / hitsDataSize); a proxy that was handed a bare pointer does not
        // know the size: pass both here
        GpuAccessProxy(int device, Codecs::AccessProxy *ap, size_t indexSize, const std::vector<std::pair<std::string, term_index_ctx>> &terms, isrc_docid_t maxDocID,
                       const uint8_t *hitsData = nullptr, size_t hitsSize = 0);
        ~GpuAccessProxy();
        GpuAccessProxy(const GpuAccessProxy &) = delete;
};

// registry the hook consults: IndexSource -> its device-resident twin (nullptr: the source stays on the CPU path)
void            gpu_proxy_register(IndexSource *src, GpuAccessProxy *gap);
GpuAccessProxy *gpu_proxy_for(IndexSource *src);

// exec_query()'s span factory for sources that have a device twin; returns nullptr when the plan holds something the GPU span does
// not execute (phrases over a LUCENE source whose hits.data was not uploaded) or the source has no twin: the caller then builds the reference's own span.
std::unique_ptr<DocsSetSpan> b200_gpu_span(queryexec_ctx &rctx, const exec_node root, const uint32_t execFlags, IndexSource *idxsrc, Similarity::IndexSourceTermsScorer *scorer);

} // namespace Trinity
