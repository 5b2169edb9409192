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

// SPARSE lists (average gap >= 48: most doc deltas are 2-byte codes, so the warp-voted 4-wide step above would fail on almost every
// window): every lane walks ITS block on its own — no votes — taking up to TWO codes of 1-2 bytes out of each 32-bit window (two such
// codes always fit), a rare side path for 3-5-byte codes; the freq section (1-byte codes) goes four at a time.
template <bool MAT>
__device__ __forceinline__ void ds_google_block_sparse(uint32_t sp, uint32_t n, uint32_t prev, uint32_t last, uint32_t *outd, uint32_t *outf, unsigned long long &sumd,
                                                       unsigned long long &sumf) {
        const uint32_t nd  = n - 1u;
        uint32_t       doc = prev, i = 0;
        while (i < nd) {
                const uint32_t a  = sp & ~3u;
                const uint32_t w  = __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (sp & 3u) * 8u); // bytes sp .. sp+3
                const uint32_t b0 = w & 0xffu;
                if (b0 >= 0xc0u) { // 3..5-byte code
                        uint32_t v, len;
                        if (b0 < 0xe0u) {
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
                        doc += v;
                        sumd += doc;
                        if (MAT)
                                outd[i] = doc;
                        sp += len;
                        ++i;
                        continue;
                }
                const uint32_t two = b0 >> 7;
                uint32_t       len = 1u + two;
                doc += two ? (((b0 & 0x3fu) << 8) | __byte_perm(w, 0u, 0x4441u)) : b0;
                sumd += doc;
                if (MAT)
                        outd[i] = doc;
                ++i;
                // the second code of the window
                const uint32_t w2 = w >> (8u * len), c0 = w2 & 0xffu;
                if (i < nd && c0 < 0xc0u) {
                        const uint32_t two2 = c0 >> 7;
                        doc += two2 ? (((c0 & 0x3fu) << 8) | ((w2 >> 8) & 0xffu)) : c0;
                        sumd += doc;
                        if (MAT)
                                outd[i] = doc;
                        ++i;
                        len += 1u + two2;
                }
                sp += len;
        }
        sumd += last;
        if (MAT)
                outd[nd] = last;
        i = 0;
        while (i < n) {
                const uint32_t a = sp & ~3u;
                const uint32_t w = __funnelshift_r(lds_u32(a), lds_u32(a + 4u), (sp & 3u) * 8u);
                if ((w & 0x80808080u) == 0u && i + 4u <= n && (!MAT || (i & 3u) == 0u)) {
                        sumf += __dp4a(w, 0x01010101u, 0u);
                        if (MAT)
                                *reinterpret_cast<uint4 *>(outf + i) = make_uint4(w & 0xffu, __byte_perm(w, 0u, 0x4441u), __byte_perm(w, 0u, 0x4442u), w >> 24);
                        sp += 4u;
                        i += 4u;
                } else {
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

// unit = 32 consecutive blocks of one term, one warp per unit, units handed out with a fixed stride.  The host lays the units out
// (DecUnit); the kernel is a three-stage software pipeline per warp: unit descriptor (u+3) -> directory entries (u+2) -> bulk copy of
// the span (u+1) -> decode (u), so no load sits on the critical path of a decode.
template <bool MAT>
__global__ void __launch_bounds__(kThreads) k_decode_stream_google(DevIndex ix, const DecUnit *units, const uint64_t *out_base, uint32_t total_units, uint32_t *docids,
                                                                  uint32_t *freqs, unsigned long long *sums) {
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

        struct Dir { // this lane's block of a unit
                uint32_t off, offn, last, prev, n, cnt, ti, g0;
        };
        auto load_desc = [&](uint32_t u) {
                DecUnit D;
                D.first_entry = D.cnt = D.term_start = D.last_n = D.ti = D.g0 = D.pad0 = D.pad1 = 0;
                if (u < total_units) {
                        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(units + u));
                        const uint2 y = __ldg(reinterpret_cast<const uint2 *>(units + u) + 2);
                        D.first_entry = x.x;
                        D.cnt         = x.y;
                        D.term_start  = x.z;
                        D.last_n      = x.w;
                        D.ti          = y.x;
                        D.g0          = y.y;
                }
                return D;
        };
        auto load_dir = [&](const DecUnit &D) {
                Dir R;
                R.off = R.offn = R.last = R.prev = R.n = 0;
                R.cnt = D.cnt;
                R.ti  = D.ti;
                R.g0  = D.g0;
                if (uint32_t(lane) < D.cnt) {
                        const uint32_t e = D.first_entry + uint32_t(lane);
                        R.off  = __ldg(ix.blk_off + e);
                        R.offn = __ldg(ix.blk_off + e + 1u);
                        R.last = __ldg(ix.blk_last + e);
                        R.prev = (D.term_start && lane == 0) ? 0u : __ldg(ix.blk_last + e - 1u);
                        R.n    = (D.last_n && uint32_t(lane) + 1u == D.cnt) ? D.last_n : bd;
                }
                return R;
        };
        uint32_t seq_issue = 0, seq_wait = 0;
        // the span of a unit: [first_off, end) with a 16-byte aligned window around it; `fits`: it goes through the staging buffer
        auto span_of = [&](const Dir &R, uint32_t &abase, uint32_t &bytes) {
                const uint32_t first_off = __shfl_sync(0xffffffffu, R.off, 0);
                const uint32_t end       = __shfl_sync(0xffffffffu, R.offn, int(max(R.cnt, 1u)) - 1);
                abase                    = first_off & ~15u;
                bytes                    = ((end + 15u) & ~15u) - abase;
                return R.cnt != 0u && bytes <= kDsStage;
        };
        auto issue = [&](const Dir &R) {
                uint32_t abase, bytes;
                if (span_of(R, abase, bytes)) {
                        if (lane == 0) {
                                const uint32_t bsel = seq_issue & 1u;
                                mbar_expect_tx(bar_s + bsel * 8u, bytes);
                                bulk_g2s(stage_s + bsel * kDsStage, ix.index + abase, bytes, bar_s + bsel * 8u);
                        }
                        ++seq_issue;
                }
        };

        uint32_t unit = blockIdx.x * kWarps + warp;
        Dir      cur  = load_dir(load_desc(unit));
        Dir      nxt  = load_dir(load_desc(unit + stride));
        DecUnit  d2   = load_desc(unit + 2u * stride);
        issue(cur);
        for (; unit < total_units; unit += stride) {
                issue(nxt);
                const Dir     nn = load_dir(d2);                    // directory loads of u+2: in flight while u is decoded
                const DecUnit d3 = load_desc(unit + 3u * stride);
                uint32_t       abase, bytes;
                const bool     staged = span_of(cur, abase, bytes);
                uint32_t       bsel   = 0;
                if (staged) {
                        bsel = seq_wait & 1u;
                        mbar_wait(bar_s + bsel * 8u, (seq_wait >> 1) & 1u);
                        ++seq_wait;
                }
                const bool         active = uint32_t(lane) < cur.cnt;
                const unsigned     m      = __ballot_sync(0xffffffffu, active);
                unsigned long long sumd = 0, sumf = 0;
                // dense or sparse walk: decided per unit from its docID span (uniform across the warp)
                const uint32_t lastDoc = __shfl_sync(0xffffffffu, cur.last, int(max(cur.cnt, 1u)) - 1), firstPrev = __shfl_sync(0xffffffffu, cur.prev, 0);
                const bool     dense   = (lastDoc - firstPrev) < 48u * cur.cnt * bd;
                if (active) {
                        const size_t row = MAT ? size_t(out_base[cur.ti]) + (size_t(cur.g0) + size_t(lane)) * bd : 0;
                        uint32_t *   od  = MAT ? docids + row : nullptr;
                        uint32_t *   of  = MAT ? freqs + row : nullptr;
                        if (!staged)
                                ds_google_block_global<MAT>(ix.index + cur.off, cur.n, cur.prev, cur.last, od, of, sumd, sumf);
                        else if (dense)
                                ds_google_block<MAT>(m, stage_s + bsel * kDsStage + (cur.off - abase), cur.n, cur.prev, cur.last, od, of, sumd, sumf);
                        else
                                ds_google_block_sparse<MAT>(stage_s + bsel * kDsStage + (cur.off - abase), cur.n, cur.prev, cur.last, od, of, sumd, sumf);
                }
                // per-term checksums (a unit never spans two terms)
                for (int d = 16; d > 0; d >>= 1) {
                        sumd += __shfl_xor_sync(0xffffffffu, sumd, d);
                        sumf += __shfl_xor_sync(0xffffffffu, sumf, d);
                }
                if (lane == 0 && sums && m) {
                        atomicAdd(&sums[2 * cur.ti], sumd);
                        atomicAdd(&sums[2 * cur.ti + 1], sumf);
                }
                __syncwarp();
                cur = nxt;
                nxt = nn;
                d2  = d3;
        }
}

// unit = 32 consecutive blocks of one term (lane j loads the directory entries of block j), the warp decodes them one after the other
template <bool MAT>
__global__ void __launch_bounds__(kThreads) k_decode_stream_lucene(DevIndex ix, const DecUnit *units, const uint64_t *out_base, uint32_t total_units, uint32_t *docids,
                                                                  uint32_t *freqs, unsigned long long *sums) {
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
        uint32_t seq_issue = 0, seq_wait = 0;
        for (uint32_t unit = blockIdx.x * kWarps + warp; unit < total_units; unit += gridDim.x * kWarps) {
                const uint4    dx   = __ldg(reinterpret_cast<const uint4 *>(units + unit));
                const uint2    dy   = __ldg(reinterpret_cast<const uint2 *>(units + unit) + 2);
                const uint32_t tcur = dy.x, g0 = dy.y, ucnt = dx.y, tailDocs = dx.w; // tailDocs: documents of the unit's last block if that is the term's tail block
                const bool     have = uint32_t(lane) < ucnt;
                uint32_t       off = 0, offn = 0, prev = 0;
                if (have) {
                        const uint32_t e = dx.x + uint32_t(lane);
                        off  = __ldg(ix.blk_off + e);
                        offn = __ldg(ix.blk_off + e + 1u);
                        prev = (dx.z && lane == 0) ? 0u : __ldg(ix.blk_last + e - 1u);
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
                        if (!(tailDocs && j + 1u == nblk)) {
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
                                uint32_t       doc  = pj;
                                for (uint32_t i = 0; i < tailDocs; ++i) {
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

cudaError_t launch_decode_stream(const DevIndex &ix, const DecUnit *units, const uint64_t *out_base, uint32_t total_units, uint32_t *docids, uint32_t *freqs,
                                 unsigned long long *sums, int num_sms, cudaStream_t stream) {
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
        void *args[] = {(void *)&ix, (void *)&units, (void *)&out_base, (void *)&total_units, (void *)&docids, (void *)&freqs, (void *)&sums};
        return cudaLaunchKernel(fn, dim3(grid), dim3(kThreads), args, smem, stream);
}
