// docID -> block lookup over the load-time directory, shared by the kernels and the host (tests pin it on the CPU).
// == skiplist_search + the header hops of Decoder::advance (google_codec.cpp:464-495,821-934; lucene_codec.cpp:596-656).
#pragma once
#include "varbyte.h"
#include <cstdint>

namespace trn {

#if defined(__CUDA_ARCH__)
#define TRN_LDG(p) __ldg(p)
#else
#define TRN_LDG(p) (*(p))
#endif

// First block of a term (nblocks > 0) whose last docID is >= d, or nblocks when there is none.  `blk_last`: the term's entries;
// `tile_first`: the term's sparse docID -> block table (codecs.h), entry j = first block whose last docID >= (tf_base + j) << tf_shift;
// tf_shift == 32: no table.  Two neighbouring table entries bound a binary search over blk_last; a table boundary needs no search.
TRN_HD uint32_t dir_first_block_ge(const uint32_t *blk_last, const uint32_t *tile_first, uint32_t nblocks, uint32_t first_doc, uint32_t last_doc, uint32_t tf_base,
                                   uint32_t tf_shift, uint32_t d) {
        if (d <= first_doc)
                return 0u;
        if (d > last_doc)
                return nblocks;
        uint32_t lo = 0u, hi = nblocks - 1u; // blk_last[nblocks - 1] == last_doc >= d
        if (tf_shift < 32u) {
                const uint32_t *tf = tile_first + ((d >> tf_shift) - tf_base);
                lo                 = TRN_LDG(tf);
                if ((d & ((1u << tf_shift) - 1u)) == 0u)
                        return lo;
                const uint32_t up = TRN_LDG(tf + 1);
                hi                = up < hi ? up : hi;
        }
        while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (TRN_LDG(blk_last + mid) < d) lo = mid + 1u;
                else hi = mid;
        }
        return lo;
}

} // namespace trn
