// Flat plans (a conjunction or a disjunction whose operands are all terms — the ENT::matchallterms / ENT::matchanyterms runs of
// compilation_ctx.h:8-30, i.e. the 2-term AND and k-term OR workloads) in k_exec_docs.  (Included by exec_docs.cuh.)
//
// The per-tile blocks of ALL operand terms form one flat (term, block) list that the warp consumes 32 at a time, each lane decoding
// its block into the bitmap of ITS term; the conjunction is one word-wise AND at the end.  Compared with running the terms one after
// another this keeps the lanes full when a term has fewer than 32 blocks in the tile (ncu on the term-at-a-time path: 11.7 of 32
// lanes active per instruction, profiles/r01_b_*).  When the rarest term is sparse inside the tile the caller falls back to the
// sequential path, whose advance()-style block skipping then saves more than the lane packing gains.
#pragma once

static constexpr uint32_t kFlatMaxLeaves = 16;

// returns 0 = not applicable (use the step program), 1 = handled (root docset in slot Q.root_slot), 2 = handled, result empty
__device__ int flat_exec_google(const ExecParams &P, const DevQuery &Q, uint32_t tile, uint32_t lo, uint32_t W, uint32_t NW, uint32_t fs, uint32_t *slots,
                                uint8_t *stage, int lane) {
        const bool isAnd = Q.flat == 1u;
        // lane j adopts the j-th leaf of the plan
        uint32_t nleaf = 0, myTerm = kEmptyTerm;
        for (uint32_t si = 0; si < Q.nsteps; ++si) {
                const DevStep st = P.steps[Q.step_begin + si];
                if (st.op == OP_LEAF) {
                        if (uint32_t(lane) == nleaf)
                                myTerm = st.term;
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
                if (T.nblocks) {
                        const uint32_t *tf = P.ix.tile_first + size_t(myTerm) * (P.ix.ntiles + 1);
                        const uint32_t  a  = tf[min(tile << fs, P.ix.ntiles)];
                        const uint32_t  b  = min(tf[min((tile + 1u) << fs, P.ix.ntiles)], T.nblocks - 1u);
                        if (a < T.nblocks && a <= b) {
                                mybA  = a;
                                mycnt = b - a + 1u;
                        }
                }
        }
        const uint32_t incl  = warp_incl_scan(uint32_t(lane) < nleaf ? mycnt : 0u, lane);
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (isAnd) {
                if (__ballot_sync(0xffffffffu, uint32_t(lane) < nleaf && mycnt == 0u))
                        return 2; // an operand has no posting in this tile
                // rarest term (operands are sorted by df) sparse in this tile => skipping beats packing
                const uint32_t cnt0 = __shfl_sync(0xffffffffu, mycnt, 0);
                if (64u * cnt0 < total - cnt0)
                        return 0;
        } else if (total == 0u)
                return 2;

        uint32_t *     root   = slots + size_t(Q.root_slot) * NW;
        const uint32_t nclear = isAnd ? nleaf : 1u;
        for (uint32_t i = lane; i < nclear * NW; i += 32)
                (isAnd ? slots : root)[i] = 0;
        __syncwarp();

        for (uint32_t g = 0; g < total; g += 32u) {
                const uint32_t f      = g + uint32_t(lane);
                const bool     active = f < total;
                uint32_t       j      = 0;
                for (uint32_t k = 0; k + 1u < nleaf; ++k)
                        j += (f >= __shfl_sync(0xffffffffu, incl, int(k))) ? 1u : 0u;
                if (!active)
                        j = nleaf - 1u;
                const uint32_t jincl = __shfl_sync(0xffffffffu, incl, int(j)), jcnt = __shfl_sync(0xffffffffu, mycnt, int(j));
                const uint32_t b     = __shfl_sync(0xffffffffu, mybA, int(j)) + (f - (jincl - jcnt));
                const uint32_t dir   = __shfl_sync(0xffffffffu, mydir, int(j));
                const uint32_t nb    = __shfl_sync(0xffffffffu, mynb, int(j));
                const uint32_t docs  = __shfl_sync(0xffffffffu, mydocs, int(j));
                uint32_t       off = 0, offn = 0, last = 0, prev = 0, n = 0;
                if (active) {
                        const uint32_t *bl = P.ix.blk_last + dir, *bo = P.ix.blk_off + dir;
                        off  = bo[b];
                        offn = bo[b + 1];
                        last = bl[b];
                        prev = b ? bl[b - 1] : 0u;
                        n    = (b + 1u == nb) ? (docs - 32u * (nb - 1u)) : 32u;
                }
                // stage the byte span of every term present in this group (a term's blocks are contiguous in the index)
                const uint32_t jmin  = __shfl_sync(0xffffffffu, j, 0);
                const uint32_t jmax  = __shfl_sync(0xffffffffu, j, int(min(31u, total - g - 1u)));
                uint32_t       sbase = 0;
                const uint8_t *p     = nullptr;
                bool           direct = false;
                for (uint32_t jj = jmin; jj <= jmax; ++jj) {
                        const uint32_t m = __ballot_sync(0xffffffffu, active && j == jj);
                        if (!m)
                                continue;
                        const int      l0 = __ffs(int(m)) - 1, l1 = 31 - __clz(int(m));
                        const uint32_t first = __shfl_sync(0xffffffffu, off, l0), end = __shfl_sync(0xffffffffu, offn, l1);
                        const uint32_t span = end - first, copied = ((first + span + 15u) & ~15u) - (first & ~15u);
                        if (sbase + copied + 32u <= kStageBytes) {
                                const uint32_t skew = stage_copy(P.ix.index, first, span, stage + sbase, lane);
                                if (active && j == jj)
                                        p = stage + sbase + skew + (off - first);
                                sbase += copied;
                        } else if (active && j == jj)
                                direct = true;
                }
                __syncwarp();
                const unsigned smemMask = __ballot_sync(0xffffffffu, active && !direct);
                if (active && !direct) {
                        BitAcc bs;
                        bs.init(isAnd ? slots + size_t(j) * NW : root);
                        google_block_docs_smem(smemMask, p, n, prev, last, lo, W, bs);
                        bs.flush();
                } else if (active) {
                        BitSink bs;
                        bs.init(isAnd ? slots + size_t(j) * NW : root, nullptr, M_OR);
                        google_block_docs<false>(P.ix.index + off, n, prev, last, lo, lo + W, bs);
                        bs.flush();
                }
                __syncwarp();
        }
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
