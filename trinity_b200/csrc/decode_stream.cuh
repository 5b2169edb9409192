// k_decode_stream_google / k_decode_stream_lucene — whole-list decode (BASELINE.json configs[4]: the postings-decode microbench; also the
// parity probe of the decoders) == PostingsListIterator::next() over whole lists (google_codec.cpp:777-819 + unpack_block :596-639;
// lucene_codec.cpp:568-594 + refill_documents :515-558 + FastPFor<4> __decodeArray fastpfor.h:222-270).  (Included by kernels.cu.)
//
// The round-1 decoders were instruction-bound at 16-24 % of the HBM roofline (profiles/r01_g, r01_j): a register-staged copy loop
// (LDG -> STS: the top stall line), a byte-wise varbyte walk through generic pointers and a second pass to find the freq section.
// Here
//   * the bytes of a unit (32 consecutive blocks of one term: a contiguous span of the chunk, positions included) arrive by ONE 1-D bulk
//     copy (cp.async.bulk.shared::cluster.global + mbarrier complete_tx, issued by one lane; SASS UBLKCP / SYNCS), double-buffered per warp:
//     the copy of unit u+1 is in flight while unit u is decoded, and no thread spends an instruction on moving bytes;
//   * GOOGLE: lane = block, all lanes in lockstep over 32-bit windows of shared memory: four 1-byte codes per step when every lane has
//     them (warp vote; dense lists and every freq section), else ONE branch-free code of 1-2 bytes per step (3-5-byte codes take a rare
//     side path); the freq section starts where the delta walk ends — one pass;
//   * LUCENE: one warp per 128-document block, vertical PFor unpack (lucene_intblock_v), the directory entries of 32 consecutive blocks
//     loaded by the 32 lanes at once.
// Every docID and freq is produced in a register (the checksum variant adds them up, the materialising variant stores them).
#pragma once

static constexpr uint32_t kDsStage      = 6144; // bytes of one staging buffer (GOOGLE: the span of 32 blocks; larger spans read global memory)
static constexpr uint32_t kDsWarpBytes  = 2 * kDsStage;
static constexpr uint32_t kDlWarpBytes  = 2 * kSfStage + kSfScratch; // LUCENE: two block buffers + exception scratch

// 64-bit accumulate of a 32-bit value (two instructions; the sums of docIDs exceed 32 bits)
__device__ __forceinline__ void acc64(unsigned long long &s, uint32_t v) {
        s += v;
}

// One lane decodes ITS block from the staged span: nd = n-1 doc deltas, then n freqs (google_codec.cpp:596-639).  `sp`: shared-space
// address of the first delta byte.  All lanes of `m` call this together (votes).  MAT: row pointers of 32 entries (16-byte aligned).
template <bool MAT>
__device__ __forceinline__ void ds_google_block(unsigned m, uint32_t sp, uint32_t n, uint32_t prev, uint32_t last, uint32_t *outd, uint32_t *outf,
                                                unsigned long long &sumd, unsigned long long &sumf) {
        const uint32_t nd  = n - 1u;
        uint32_t       doc = prev, i = 0;
        uint4          rowbuf = make_uint4(0, 0, 0, 0);
        // ---- doc deltas
        for (;;) {
                const bool live = i < nd;
                if (!__any_sync(m, live))
                        break;
                const uint32_t a = sp & ~3u;
                const uint32_t w = __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (sp & 3u) * 8u); // bytes sp .. sp+3
                const bool     f4 = (w & 0x80808080u) == 0u && i + 4u <= nd && (!MAT || (i & 3u) == 0u);
                if (__all_sync(m, !live || f4)) {
                        if (live) {
                                const uint32_t d0 = doc + (w & 0xffu), d1 = d0 + __byte_perm(w, 0u, 0x4441u), d2 = d1 + __byte_perm(w, 0u, 0x4442u), d3 = d2 + (w >> 24);
                                doc = d3;
                                sumd += static_cast<unsigned long long>(d0) + d1 + d2 + d3;
                                if (MAT)
                                        *reinterpret_cast<uint4 *>(outd + i) = make_uint4(d0, d1, d2, d3);
                                sp += 4u;
                                i += 4u;
                        }
                } else if (live) {
                        const uint32_t b0 = w & 0xffu;
                        uint32_t       v, len;
                        if (b0 < 0xc0u) { // 1- or 2-byte code, branch-free
                                const uint32_t two = b0 >> 7;
                                v   = two ? (((b0 & 0x3fu) << 8) | __byte_perm(w, 0u, 0x4441u)) : b0;
                                len = 1u + two;
                        } else if (b0 < 0xe0u) {
                                v   = ((b0 & 0x1fu) << 16) | ((w >> 8) & 0xffffu);
                                len = 3u;
                        } else if (b0 < 0xf0u) {
                                v   = ((b0 & 0x0fu) << 24) | (((w >> 8) & 0xffu) << 16) | (((w >> 16) & 0xffu) << 8) | (w >> 24);
                                len = 4u;
                        } else { // u32le in bytes 1..4
                                const uint32_t a1 = (sp + 1u) & ~3u;
                                v   = __funnelshift_r(lds_u32(a1), lds_u32(a1 + 4u), ((sp + 1u) & 3u) * 8u);
                                len = 5u;
                        }
                        doc += v;
                        acc64(sumd, doc);
                        if (MAT) {
                                const uint32_t s4 = i & 3u;
                                if (s4 == 0u) rowbuf.x = doc;
                                else if (s4 == 1u) rowbuf.y = doc;
                                else if (s4 == 2u) rowbuf.z = doc;
                                else {
                                        rowbuf.w = doc;
                                        *reinterpret_cast<uint4 *>(outd + (i & ~3u)) = rowbuf;
                                }
                        }
                        sp += len;
                        ++i;
                }
        }
        // the block's last document is implied by the header / the directory
        acc64(sumd, last);
        if (MAT) {
                const uint32_t s4 = nd & 3u, base = nd & ~3u;
                if (s4 == 0u) outd[base] = last;
                else if (s4 == 1u) { outd[base] = rowbuf.x; outd[base + 1] = last; }
                else if (s4 == 2u) { outd[base] = rowbuf.x; outd[base + 1] = rowbuf.y; outd[base + 2] = last; }
                else *reinterpret_cast<uint4 *>(outd + base) = make_uint4(rowbuf.x, rowbuf.y, rowbuf.z, last);
        }
        // ---- freqs (the section starts where the delta walk ended)
        i = 0;
        for (;;) {
                const bool live = i < n;
                if (!__any_sync(m, live))
                        break;
                const uint32_t a = sp & ~3u;
                const uint32_t w = __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (sp & 3u) * 8u);
                const bool     f4 = (w & 0x80808080u) == 0u && i + 4u <= n && (!MAT || (i & 3u) == 0u);
                if (__all_sync(m, !live || f4)) {
                        if (live) {
                                sumf += __dp4a(w, 0x01010101u, 0u);
                                if (MAT)
                                        *reinterpret_cast<uint4 *>(outf + i) = make_uint4(w & 0xffu, __byte_perm(w, 0u, 0x4441u), __byte_perm(w, 0u, 0x4442u), w >> 24);
                                sp += 4u;
                                i += 4u;
                        }
                } else if (live) {
                        const uint32_t b0 = w & 0xffu;
                        uint32_t       v, len;
                        if (b0 < 0xc0u) {
                                const uint32_t two = b0 >> 7;
                                v   = two ? (((b0 & 0x3fu) << 8) | __byte_perm(w, 0u, 0x4441u)) : b0;
                                len = 1u + two;
                        } else if (b0 < 0xe0u) {
                                v   = ((b0 & 0x1fu) << 16) | ((w >> 8) & 0xffffu);
                                len = 3u;
                        } else if (b0 < 0xf0u) {
                                v   = ((b0 & 0x0fu) << 24) | (((w >> 8) & 0xffu) << 16) | (((w >> 16) & 0xffu) << 8) | (w >> 24);
                                len = 4u;
                        } else {
                                const uint32_t a1 = (sp + 1u) & ~3u;
                                v   = __funnelshift_r(lds_u32(a1), lds_u32(a1 + 4u), ((sp + 1u) & 3u) * 8u);
                                len = 5u;
                        }
                        sumf += v;
                        if (MAT)
                                outf[i] = v;
                        sp += len;
                        ++i;
                }
        }
}

// generic-pointer form of the same walk for spans that do not fit the staging buffer (positions-heavy blocks): straight from global memory
template <bool MAT>
__device__ __forceinline__ void ds_google_block_global(const uint8_t *p, uint32_t n, uint32_t prev, uint32_t last, uint32_t *outd, uint32_t *outf,
                                                       unsigned long long &sumd, unsigned long long &sumf) {
        uint32_t doc = prev;
        for (uint32_t i = 0; i + 1u < n; ++i) {
                doc += varbyte_get(p);
                acc64(sumd, doc);
                if (MAT)
                        outd[i] = doc;
        }
        acc64(sumd, last);
        if (MAT)
                outd[n - 1u] = last;
        for (uint32_t i = 0; i < n; ++i) {
                const uint32_t v = varbyte_get(p);
                sumf += v;
                if (MAT)
                        outf[i] = v;
        }
}

// unit = 32 consecutive blocks of one term, one warp per unit, units handed out with a fixed stride
template <bool MAT>
__global__ void __launch_bounds__(kThreads) k_decode_stream_google(DevIndex ix, const uint32_t *term_ids, const uint32_t *unit_base /*nterms+1*/, const uint64_t *out_base,
                                                                  uint32_t nterms, uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums) {
        __shared__ __align__(8) unsigned long long s_bar[kWarps * 2];
        const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint8_t *      stage   = dyn_smem + size_t(warp) * kDsWarpBytes;
        const uint32_t stage_s = uint32_t(__cvta_generic_to_shared(stage));
        const uint32_t bar_s   = uint32_t(__cvta_generic_to_shared(&s_bar[warp * 2]));
        if (lane == 0) {
                mbar_init(bar_s, 1);
                mbar_init(bar_s + 8, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const uint32_t stride = gridDim.x * kWarps;
        const uint32_t bd     = ix.block_docs;

        struct Unit {
                uint32_t ti, off, n, prev, last, b, first_off, span;
                bool     active;
        };
        uint32_t tcur = 0; // units are visited in ascending order: the term index only moves forward
        auto locate = [&](uint32_t unit) {
                Unit U;
                U.active = false;
                U.ti = U.off = U.n = U.prev = U.last = U.b = U.first_off = U.span = 0;
                if (unit >= total_units)
                        return U;
                while (tcur + 1u < nterms && __ldg(unit_base + tcur + 1u) <= unit)
                        ++tcur;
                const DevTerm  T  = ix.terms[__ldg(term_ids + tcur)];
                const uint32_t g0 = (unit - __ldg(unit_base + tcur)) * 32u, b = g0 + uint32_t(lane);
                U.ti              = tcur;
                U.b               = b;
                uint32_t offn     = 0;
                if (b < T.nblocks) {
                        const uint32_t *bl = ix.blk_last + T.dir_begin, *bo = ix.blk_off + T.dir_begin;
                        U.active = true;
                        U.off    = __ldg(bo + b);
                        offn     = __ldg(bo + b + 1u);
                        U.last   = __ldg(bl + b);
                        U.prev   = b ? __ldg(bl + b - 1u) : 0u;
                        U.n      = (b + 1u == T.nblocks) ? (T.documents - bd * (T.nblocks - 1u)) : bd;
                }
                const uint32_t cnt = min(32u, T.nblocks - min(T.nblocks, g0));
                U.first_off        = __shfl_sync(0xffffffffu, U.off, 0);
                const uint32_t end = __shfl_sync(0xffffffffu, offn, int(max(cnt, 1u)) - 1);
                U.span             = cnt ? end - U.first_off : 0u;
                return U;
        };
        uint32_t seq_issue = 0, seq_wait = 0;
        auto     issue     = [&](const Unit &U) { // one bulk copy of the unit's span (only when it fits: otherwise the lanes read global memory)
                const uint32_t abase = U.first_off & ~15u, bytes = ((U.first_off + U.span + 15u) & ~15u) - abase;
                if (U.span && bytes <= kDsStage) {
                        if (lane == 0) {
                                const uint32_t bsel = seq_issue & 1u;
                                mbar_expect_tx(bar_s + bsel * 8u, bytes);
                                bulk_g2s(stage_s + bsel * kDsStage, ix.index + abase, bytes, bar_s + bsel * 8u);
                        }
                        ++seq_issue;
                }
        };

        uint32_t unit = blockIdx.x * kWarps + warp;
        Unit     cur  = locate(unit);
        Unit     nxt  = locate(unit + stride);
        issue(cur);
        for (; unit < total_units; unit += stride) {
                issue(nxt);
                const Unit nn = locate(unit + 2u * stride); // its directory loads are in flight while this unit is decoded
                const uint32_t abase = cur.first_off & ~15u, bytes = ((cur.first_off + cur.span + 15u) & ~15u) - abase;
                const bool     staged = cur.span && bytes <= kDsStage;
                uint32_t       bsel   = 0;
                if (staged) {
                        bsel = seq_wait & 1u;
                        mbar_wait(bar_s + bsel * 8u, (seq_wait >> 1) & 1u);
                        ++seq_wait;
                }
                const unsigned     m    = __ballot_sync(0xffffffffu, cur.active);
                unsigned long long sumd = 0, sumf = 0;
                if (cur.active) {
                        const size_t row = MAT ? size_t(out_base[cur.ti]) + size_t(cur.b) * bd : 0;
                        uint32_t *   od  = MAT ? docids + row : nullptr;
                        uint32_t *   of  = MAT ? freqs + row : nullptr;
                        if (staged)
                                ds_google_block<MAT>(m, stage_s + bsel * kDsStage + (cur.off - abase), cur.n, cur.prev, cur.last, od, of, sumd, sumf);
                        else
                                ds_google_block_global<MAT>(ix.index + cur.off, cur.n, cur.prev, cur.last, od, of, sumd, sumf);
                }
                // per-term checksums (a unit never spans two terms)
                for (int d = 16; d > 0; d >>= 1) {
                        sumd += __shfl_xor_sync(0xffffffffu, sumd, d);
                        sumf += __shfl_xor_sync(0xffffffffu, sumf, d);
                }
                const uint32_t ti = __shfl_sync(0xffffffffu, cur.ti, 0);
                if (lane == 0 && sums && m) {
                        atomicAdd(&sums[2 * ti], sumd);
                        atomicAdd(&sums[2 * ti + 1], sumf);
                }
                __syncwarp();
                cur = nxt;
                nxt = nn;
        }
}

// unit = 32 consecutive blocks of one term (lane j loads the directory entries of block j), the warp decodes them one after the other
template <bool MAT>
__global__ void __launch_bounds__(kThreads) k_decode_stream_lucene(DevIndex ix, const uint32_t *term_ids, const uint32_t *unit_base /*nterms+1*/, const uint64_t *out_base,
                                                                  uint32_t nterms, uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums) {
        __shared__ __align__(8) unsigned long long s_bar[kWarps * 2];
        const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint8_t *      stage   = dyn_smem + size_t(warp) * kDlWarpBytes;
        uint32_t *     scratch = reinterpret_cast<uint32_t *>(stage + 2 * kSfStage);
        const uint32_t stage_s = uint32_t(__cvta_generic_to_shared(stage));
        const uint32_t bar_s   = uint32_t(__cvta_generic_to_shared(&s_bar[warp * 2]));
        if (lane == 0) {
                mbar_init(bar_s, 1);
                mbar_init(bar_s + 8, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        uint32_t seq_issue = 0, seq_wait = 0, tcur = 0;
        for (uint32_t unit = blockIdx.x * kWarps + warp; unit < total_units; unit += gridDim.x * kWarps) {
                while (tcur + 1u < nterms && __ldg(unit_base + tcur + 1u) <= unit)
                        ++tcur;
                const DevTerm  T     = ix.terms[__ldg(term_ids + tcur)];
                const uint32_t g0    = (unit - __ldg(unit_base + tcur)) * 32u, b = g0 + uint32_t(lane);
                const uint32_t nfull = T.documents >> 7;
                const bool     have  = b < T.nblocks;
                uint32_t       off = 0, offn = 0, prev = 0;
                if (have) {
                        off  = __ldg(ix.blk_off + T.dir_begin + b);
                        offn = __ldg(ix.blk_off + T.dir_begin + b + 1u);
                        prev = b ? __ldg(ix.blk_last + T.dir_begin + b - 1u) : 0u;
                }
                const uint32_t nblk = __popc(__ballot_sync(0xffffffffu, have));
                auto           issue = [&](uint32_t j) {
                        const uint32_t o = __shfl_sync(0xffffffffu, off, int(j)), on = __shfl_sync(0xffffffffu, offn, int(j));
                        const uint32_t abase = o & ~15u, bytes = min(((on + 15u) & ~15u) - abase, kSfStage);
                        if (lane == 0) {
                                const uint32_t bsel = seq_issue & 1u;
                                mbar_expect_tx(bar_s + bsel * 8u, bytes);
                                bulk_g2s(stage_s + bsel * kSfStage, ix.index + abase, bytes, bar_s + bsel * 8u);
                        }
                        ++seq_issue;
                };
                unsigned long long sumd = 0, sumf = 0;
                const size_t       trow = MAT ? size_t(out_base[tcur]) : 0;
                if (nblk)
                        issue(0);
                for (uint32_t j = 0; j < nblk; ++j) {
                        if (j + 1u < nblk)
                                issue(j + 1u);
                        const uint32_t bsel = seq_wait & 1u;
                        mbar_wait(bar_s + bsel * 8u, (seq_wait >> 1) & 1u);
                        ++seq_wait;
                        const uint8_t *s    = stage + bsel * kSfStage;
                        const uint32_t oj   = __shfl_sync(0xffffffffu, off, int(j));
                        const uint32_t pj   = __shfl_sync(0xffffffffu, prev, int(j));
                        const uint32_t bj   = g0 + j;
                        const uint32_t skew = oj & 15u;
                        uint32_t *     od   = MAT ? docids + trow + size_t(bj) * 128u : nullptr;
                        uint32_t *     of   = MAT ? freqs + trow + size_t(bj) * 128u : nullptr;
                        if (bj < nfull) {
                                uint32_t d[4], fr[4], dbits, fbits;
                                const uint32_t o2 = lucene_intblock_v(s, skew, lane, d, scratch, dbits);
                                (void)lucene_intblock_v(s, o2, lane, fr, scratch, fbits);
                                uint32_t base = pj;
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                        const uint32_t sc = warp_incl_scan(d[g], lane);
                                        d[g]              = base + sc;
                                        base += __shfl_sync(0xffffffffu, sc, 31);
                                        sumd += d[g];
                                        sumf += fr[g];
                                        if (MAT) { // value l + 32 g: consecutive lanes store consecutive words
                                                od[lane + 32 * g] = d[g];
                                                of[lane + 32 * g] = fr[g];
                                        }
                                }
                        } else if (lane == 0) { // tail block: (varbyte delta, varbyte freq) pairs (lucene_codec.cpp:527-550)
                                const uint8_t *pp   = s + skew;
                                const uint32_t tail = T.documents & 127u;
                                uint32_t       doc  = pj;
                                for (uint32_t i = 0; i < tail; ++i) {
                                        doc += varbyte_get(pp);
                                        const uint32_t f = varbyte_get(pp);
                                        sumd += doc;
                                        sumf += f;
                                        if (MAT) {
                                                od[i] = doc;
                                                of[i] = f;
                                        }
                                }
                        }
                        __syncwarp();
                }
                for (int d = 16; d > 0; d >>= 1) {
                        sumd += __shfl_xor_sync(0xffffffffu, sumd, d);
                        sumf += __shfl_xor_sync(0xffffffffu, sumf, d);
                }
                if (lane == 0 && sums && nblk) {
                        atomicAdd(&sums[2 * tcur], sumd);
                        atomicAdd(&sums[2 * tcur + 1], sumf);
                }
                __syncwarp();
        }
}

cudaError_t launch_decode_stream(const DevIndex &ix, const uint32_t *term_ids, const uint32_t *unit_base, const uint64_t *out_base, uint32_t nterms,
                                 uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums, int num_sms, cudaStream_t stream) {
        const bool   mat  = docids != nullptr;
        const size_t smem = size_t(kWarps) * (ix.codec == 0 ? kDsWarpBytes : kDlWarpBytes);
        const void * fn   = ix.codec == 0 ? (mat ? (const void *)k_decode_stream_google<true> : (const void *)k_decode_stream_google<false>)
                                          : (mat ? (const void *)k_decode_stream_lucene<true> : (const void *)k_decode_stream_lucene<false>);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess)
                return e;
        int per = 0;
        e       = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, kThreads, smem);
        if (e != cudaSuccess)
                return e;
        const int grid = int(std::min<uint64_t>(uint64_t(num_sms) * std::max(per, 1), (uint64_t(total_units) + kWarps - 1) / kWarps));
        void *args[] = {(void *)&ix, (void *)&term_ids, (void *)&unit_base, (void *)&out_base, (void *)&nterms, (void *)&total_units, (void *)&docids, (void *)&freqs, (void *)&sums};
        return cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), args, smem, stream);
}
