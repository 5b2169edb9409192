// Flat plans (a conjunction or a disjunction whose operands are all terms — the ENT::matchallterms / ENT::matchanyterms runs of
// compilation_ctx.h:8-30, i.e. the 2-term AND and k-term OR workloads) in k_exec_docs.  (Included by exec_docs.cuh.)
//
// The per-tile blocks of ALL operand terms form one flat (term, block) list that the warp consumes 32 at a time, each lane decoding
// its block into the bitmap of ITS term; the conjunction is one word-wise AND at the end.  Compared with running the terms one after
// another this keeps the lanes full when a term has fewer than 32 blocks in the tile (ncu on the term-at-a-time path: 11.7 of 32
// lanes active per instruction, profiles/r01_b_*).  When the rarest term is sparse inside the tile the caller falls back to the
// sequential path, whose advance()-style block skipping then saves more than the lane packing gains.
//
// Staging: every lane copies only the head of ITS block (kGatherBytes from the 16B-aligned address below the first doc-delta byte)
// with cp.async (LDGSTS) into its own slot — the doc-delta section of a 32-doc block is at most 31 x 2 bytes for gaps < 16384, and
// the inline hits behind it (half of the index bytes) never enter the SM.  (The span-copy version kept 6 KB per warp for one group,
// capped the SM at 20 resident warps and exposed every group's global-load latency: 14 % of all stall samples, profiles/r01_c_*.)
#pragma once

static constexpr uint32_t kFlatMaxLeaves = 16;

struct FlatLane { // per-lane description of one (term, block) work unit
        uint32_t j, off, n, prev, last;
        bool     active;
};

// returns 0 = not applicable (use the step program), 1 = handled (root docset in slot Q.root_slot), 2 = handled, result empty
__device__ int flat_exec_google(const ExecParams &P, const DevQuery &Q, uint32_t lo, uint32_t W, uint32_t NW, uint32_t *slots, uint8_t *stage, int lane) {
        const bool isTree = Q.flat == 5u; // flat-tree: every leaf owns the bitmap its OP_LEAF step names; the slot program follows in the caller
        const bool isAnd  = Q.flat == 1u || isTree;
        // lane j adopts the j-th leaf of the plan
        uint32_t nleaf = 0, myTerm = kEmptyTerm, mySlot = 0;
        for (uint32_t si = 0; si < Q.nsteps; ++si) {
                const DevStep st = P.steps[Q.step_begin + si];
                if (st.op == OP_LEAF) {
                        if (uint32_t(lane) == nleaf) {
                                myTerm = st.term;
                                mySlot = isTree ? st.dst : nleaf;
                        }
                        ++nleaf;
                }
        }
        if (nleaf == 0 || nleaf > kFlatMaxLeaves || (isAnd && nleaf > P.nslots))
                return 0;
        uint32_t mybA = 0, mycnt = 0, mydir = 0, mynb = 0, mydocs = 0;
        if (uint32_t(lane) < nleaf && myTerm != kEmptyTerm) {
                const DevTerm T = P.ix.terms[myTerm];
                mydir           = T.dir_begin;
                mynb            = T.nblocks;
                mydocs          = T.documents;
                uint32_t a, b;
                tile_block_range(P.ix, T, lo, W, a, b);
                if (a <= b) {
                        mybA  = a;
                        mycnt = b - a + 1u;
                }
        }
        const uint32_t incl  = warp_incl_scan(uint32_t(lane) < nleaf ? mycnt : 0u, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (isTree) {
                // nothing to decide here: every leaf is decoded, the slot program does the rest
        } else if (isAnd) {
                if (__ballot_sync(0xffffffffu, uint32_t(lane) < nleaf && mycnt == 0u))
                        return 2; // an operand has no posting in this tile
                // rarest term (operands are sorted by df) sparse in this tile => skipping beats packing
                const uint32_t cnt0 = __shfl_sync(0xffffffffu, mycnt, 0);
                if (64u * cnt0 < total - cnt0)
                        return 0;
        } else if (total == 0u)
                return 2;

        uint32_t *     root   = slots + size_t(Q.root_slot) * NW;
        const uint32_t nclear = isAnd ? nleaf : 1u; // (flat-tree plans keep their leaves in slots 0 .. nleaf-1)
        for (uint32_t i = lane; i < nclear * NW; i += 32)
                (isAnd ? slots : root)[i] = 0;

        // lane assignment of group g
        auto assign = [&](uint32_t g) {
                FlatLane       L;
                const uint32_t f = g + uint32_t(lane);
                L.active         = f < total;
                uint32_t j       = 0;
                for (uint32_t k = 0; k + 1u < nleaf; ++k)
                        j += (f >= __shfl_sync(0xffffffffu, incl, int(k))) ? 1u : 0u;
                if (!L.active)
                        j = nleaf - 1u;
                const uint32_t jincl = __shfl_sync(0xffffffffu, incl, int(j)), jcnt = __shfl_sync(0xffffffffu, mycnt, int(j));
                const uint32_t b     = __shfl_sync(0xffffffffu, mybA, int(j)) + (f - (jincl - jcnt));
                const uint32_t dir   = __shfl_sync(0xffffffffu, mydir, int(j));
                const uint32_t nb    = __shfl_sync(0xffffffffu, mynb, int(j));
                const uint32_t docs  = __shfl_sync(0xffffffffu, mydocs, int(j));
                L.j                  = __shfl_sync(0xffffffffu, mySlot, int(j)); // the bitmap this block goes into
                L.off = L.n = L.prev = L.last = 0;
                if (L.active) {
                        const uint32_t *bl = P.ix.blk_last + dir, *bo = P.ix.blk_off + dir;
                        L.off  = bo[b];
                        L.last = bl[b];
                        L.prev = b ? bl[b - 1] : 0u;
                        L.n    = (b + 1u == nb) ? (docs - 32u * (nb - 1u)) : 32u;
                }
                return L;
        };

        // conjunctions: one bitmap per operand => plain-store word builder (OwnAcc); the last word of every block is ORed in
        // atomically one group LATER, after every block that can share it has stored its words
        const bool     own     = isAnd;
        const uint32_t dummy   = uint32_t(__cvta_generic_to_shared(stage + kGatherBufBytes)) + uint32_t(lane) * 4u;
        const uint32_t slots_s = uint32_t(__cvta_generic_to_shared(slots));
        uint32_t       tail_a = dummy, tail_bits = 0;
        FlatLane   cur  = assign(0);
        gather_issue(P.ix.index, cur.off, cur.active, stage, lane);
        __syncwarp(); // slot clears above are visible before the first reduction
        for (uint32_t g = 0; g < total; g += 32u) {
                const bool more = g + 32u < total;
                gather_wait<0>();
                const unsigned m = __ballot_sync(0xffffffffu, cur.active);
                if (own) {
                        OwnAcc bs;
                        bs.init(slots_s + cur.j * NW * 4u, dummy);
                        if (cur.active)
                                google_block_docs_vote(m, P.ix.index, cur.off, stage, lane, cur.n, cur.prev, cur.last, lo, W, bs);
                        __syncwarp();
                        if (tail_bits) // the previous group's last words
                                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(tail_a), "r"(tail_bits) : "memory");
                        tail_a    = bs.cur_a;
                        tail_bits = bs.cur;
                } else if (cur.active) {
                        BitAcc bs;
                        bs.init(root);
                        google_block_docs_gather(m, P.ix.index, cur.off, stage, lane, cur.n, cur.prev, cur.last, lo, W, bs);
                        bs.flush();
                }
                __syncwarp();
                if (more) {
                        cur = assign(g + 32u);
                        gather_issue(P.ix.index, cur.off, cur.active, stage, lane);
                }
        }
        if (own && tail_bits)
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(tail_a), "r"(tail_bits) : "memory");
        __syncwarp();
        if (isTree)
                return 0; // leaves are in place: the caller runs the slot operations of the plan
        if (isAnd) {
                // operand i lives in slot i; the root of an all-term conjunction is slot 0
                for (uint32_t i = lane; i < NW; i += 32) {
                        uint32_t w = slots[i];
                        for (uint32_t k = 1; k < nleaf; ++k)
                                w &= slots[size_t(k) * NW + i];
                        root[i] = w;
                }
                __syncwarp();
        }
        return 1;
}

// One operand of a general step program decoded with the plain-store word builder: blocks [bA, bB] of term T into `out`, a bitmap
// that only this call writes (cleared by the caller).  Same loop as the conjunction case of flat_exec_google for a single term.
// `dummy`: shared address of 32 scratch words (one per lane).
__device__ void google_leaf_own(const DevIndex &ix, const DevTerm &T, uint32_t bA, uint32_t bB, uint32_t lo, uint32_t W, uint32_t *out, uint8_t *stage,
                                uint32_t dummy, int lane) {
        const uint32_t *bl = ix.blk_last + T.dir_begin, *bo = ix.blk_off + T.dir_begin;
        const uint32_t  out_s = uint32_t(__cvta_generic_to_shared(out));
        uint32_t        tail_a = dummy, tail_bits = 0;
        for (uint32_t g = bA; g <= bB; g += 32u) {
                const uint32_t b      = g + uint32_t(lane);
                const bool     active = b <= bB;
                uint32_t       off = 0, last = 0, prev = 0, n = 0;
                if (active) {
                        off  = bo[b];
                        last = bl[b];
                        prev = b ? bl[b - 1] : 0u;
                        n    = (b + 1u == T.nblocks) ? (T.documents - 32u * (T.nblocks - 1u)) : 32u;
                }
                gather_issue(ix.index, off, active, stage, lane);
                gather_wait<0>();
                const unsigned m = __ballot_sync(0xffffffffu, active);
                OwnAcc         bs;
                bs.init(out_s, dummy);
                if (active)
                        google_block_docs_vote(m, ix.index, off, stage, lane, n, prev, last, lo, W, bs);
                __syncwarp();
                if (tail_bits) // the previous group's last words, after every block that can share them has stored
                        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(tail_a), "r"(tail_bits) : "memory");
                tail_a    = bs.cur_a;
                tail_bits = bs.cur;
        }
        if (tail_bits)
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(tail_a), "r"(tail_bits) : "memory");
        __syncwarp();
}
