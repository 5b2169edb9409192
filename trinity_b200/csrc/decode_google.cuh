// k_decode_google — whole-list decode of GOOGLE-codec terms (microbench + parity probe) == PostingsListIterator::next() over a list
// (google_codec.cpp:777-819 next, :596-639 unpack_block).  (Included by kernels.cu after exec_docs.cuh.)
//
// One warp per 32 consecutive blocks, one lane per block, all lanes in lockstep.  A block's doc-delta varbytes are immediately
// followed by its freq varbytes, so ONE pass over one byte stream yields both; the inline hits behind them are never touched (each lane
// stages only the head of its block with cp.async, double-buffered across units).  Runs of 1-byte codes are consumed four at a time
// when every lane can (warp vote).  Materialised output is written as 16-byte vectors, (docID x4) and (freq x4).
#pragma once

static constexpr uint32_t kDecGatherBytes = 112; // 7 x 16 B: <= 62 B of deltas (gaps < 16384) + 32 B of freqs + 15 B alignment slack
static constexpr uint32_t kDecGatherWords = kDecGatherBytes / 4;
static constexpr uint32_t kDecBufBytes    = 32 * kDecGatherBytes;

struct DecWords {
        const uint32_t *slot, *g32;
        uint32_t        k;
        __device__ __forceinline__ uint32_t next() {
                const uint32_t w = k < kDecGatherWords ? slot[k] : __ldg(g32 + k);
                ++k;
                return w;
        }
};

__device__ __forceinline__ void dec_gather_issue(const uint8_t *__restrict__ index, uint32_t off, bool need, uint8_t *buf, int lane) {
        if (need) {
                const uint8_t *src = index + (off & ~15u);
                const uint32_t dst = uint32_t(__cvta_generic_to_shared(buf + lane * kDecGatherBytes));
#pragma unroll
                for (uint32_t c = 0; c < kDecGatherBytes; c += 16u)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c), "l"(src + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
}

// emits value #i of a 32-entry row through a 4-entry register buffer -> one 16-byte store per 4 values
struct Row4 {
        uint32_t *out; // row base (128 B aligned) or nullptr
        uint4     buf;
        __device__ __forceinline__ void put(uint32_t i, uint32_t v) {
                const uint32_t s = i & 3u;
                if (s == 0u) buf.x = v;
                else if (s == 1u) buf.y = v;
                else if (s == 2u) buf.z = v;
                else {
                        buf.w = v;
                        if (out)
                                *reinterpret_cast<uint4 *>(out + (i & ~3u)) = buf;
                }
        }
        __device__ __forceinline__ void put4(uint32_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { // i % 4 == 0
                if (out)
                        *reinterpret_cast<uint4 *>(out + i) = make_uint4(a, b, c, d);
        }
        __device__ __forceinline__ void finish(uint32_t n) { // tail of a short (last) block
                if (!out)
                        return;
                const uint32_t r = n & 3u, base = n & ~3u;
                if (r > 0u) out[base] = buf.x;
                if (r > 1u) out[base + 1] = buf.y;
                if (r > 2u) out[base + 2] = buf.z;
        }
};

template <bool LOCKSTEP>
__global__ void __launch_bounds__(kThreads) k_decode_google(DevIndex ix, const uint32_t *term_ids, const uint32_t *unit_base /*nterms+1*/, const uint64_t *out_base,
                                                            uint32_t nterms, uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums) {
        extern __shared__ __align__(16) uint8_t smem[];
        const int                              lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint8_t *                              stage = smem + warp * (2 * kDecBufBytes);
        const uint32_t                         stride = gridDim.x * kWarps;

        struct Unit {
                uint32_t ti, off, n, prev, last, b;
                bool     active;
        };
        auto locate = [&](uint32_t unit) {
                Unit U;
                U.active = false;
                U.ti = U.off = U.n = U.prev = U.last = U.b = 0;
                if (unit >= total_units)
                        return U;
                uint32_t tlo = 0, thi = nterms;
                while (thi - tlo > 1) {
                        const uint32_t mid = (tlo + thi) >> 1;
                        if (unit_base[mid] <= unit) tlo = mid;
                        else thi = mid;
                }
                const DevTerm   T  = ix.terms[term_ids[tlo]];
                const uint32_t  b  = (unit - unit_base[tlo]) * 32u + uint32_t(lane);
                U.ti               = tlo;
                U.b                = b;
                if (b < T.nblocks) {
                        const uint32_t *bl = ix.blk_last + T.dir_begin, *bo = ix.blk_off + T.dir_begin;
                        U.active = true;
                        U.off    = bo[b];
                        U.last   = bl[b];
                        U.prev   = b ? bl[b - 1] : 0u;
                        U.n      = (b + 1u == T.nblocks) ? (T.documents - 32u * (T.nblocks - 1u)) : 32u;
                }
                return U;
        };

        uint32_t unit = blockIdx.x * kWarps + warp;
        Unit     cur  = locate(unit);
        dec_gather_issue(ix.index, cur.off, cur.active, stage, lane);
        uint32_t bufSel = 0;
        for (; unit < total_units; unit += stride) {
                Unit nxt = locate(unit + stride);
                if (unit + stride < total_units) {
                        dec_gather_issue(ix.index, nxt.off, nxt.active, stage + (bufSel ^ 1u) * kDecBufBytes, lane);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                const unsigned     m    = __ballot_sync(0xffffffffu, cur.active);
                unsigned long long sumd = 0, sumf = 0;
                if (cur.active) {
                        const uint32_t A = cur.off & ~15u, mis = cur.off - A;
                        DecWords       src{reinterpret_cast<const uint32_t *>(stage + bufSel * kDecBufBytes + lane * kDecGatherBytes),
                                     reinterpret_cast<const uint32_t *>(ix.index + A), mis >> 2};
                        const uint32_t     w0 = src.next(), w1 = src.next();
                        unsigned long long win = (static_cast<unsigned long long>(w1) << 32 | w0) >> ((mis & 3u) * 8u);
                        uint32_t           avail = 8u - (mis & 3u), nw = src.next();
                        const size_t       row = out_base[cur.ti] + size_t(cur.b) * 32u;
                        Row4               rd{docids ? docids + row : nullptr, make_uint4(0, 0, 0, 0)}, rf{freqs ? freqs + row : nullptr, make_uint4(0, 0, 0, 0)};
                        // the stream holds nd = n-1 deltas followed by n freqs: total n + nd codes; code index c < nd is a delta
                        const uint32_t nd = cur.n - 1u, ncodes = cur.n + nd;
                        uint32_t       doc = cur.prev, c = 0;
                        for (;;) {
                                const bool live = c < ncodes;
                                if (LOCKSTEP ? !__any_sync(m, live) : !live)
                                        break;
                                if (live && avail < 4u) {
                                        win |= static_cast<unsigned long long>(nw) << (avail * 8u);
                                        avail += 4u;
                                        nw = src.next();
                                }
                                const uint32_t b = uint32_t(win);
                                // four 1-byte codes that do not straddle the delta/freq boundary and start at a multiple of 4 within their row
                                const bool inDelta = c < nd;
                                const uint32_t i   = inDelta ? c : c - nd; // index within docs (for deltas) or within freqs
                                const bool fast    = (b & 0x80808080u) == 0u && (i & 3u) == 0u && (inDelta ? c + 4u <= nd : c + 4u <= ncodes);
                                if (LOCKSTEP ? __all_sync(m, !live || fast) : fast) {
                                        if (live) {
                                                const uint32_t b0 = b & 0xffu, b1 = (b >> 8) & 0xffu, b2 = (b >> 16) & 0xffu, b3 = b >> 24;
                                                if (inDelta) {
                                                        const uint32_t d0 = doc + b0, d1 = d0 + b1, d2 = d1 + b2, d3 = d2 + b3;
                                                        doc = d3;
                                                        rd.put4(i, d0, d1, d2, d3);
                                                        sumd += static_cast<unsigned long long>(d0) + d1 + d2 + d3;
                                                } else {
                                                        rf.put4(i, b0, b1, b2, b3);
                                                        sumf += b0 + b1 + b2 + b3;
                                                }
                                                win >>= 32;
                                                avail -= 4u;
                                                c += 4u;
                                        }
                                } else if (live) {
                                        uint32_t       v, len;
                                        const uint32_t b0 = b & 0xffu;
                                        if (b0 < 0x80u) {
                                                v   = b0;
                                                len = 1u;
                                        } else if (b0 < 0xc0u) {
                                                v   = ((b0 & 0x3fu) << 8) | ((b >> 8) & 0xffu);
                                                len = 2u;
                                        } else if (b0 < 0xe0u) {
                                                v   = ((b0 & 0x1fu) << 16) | ((b >> 8) & 0xffffu);
                                                len = 3u;
                                        } else if (b0 < 0xf0u) {
                                                v   = ((b0 & 0x0fu) << 24) | (((b >> 8) & 0xffu) << 16) | (((b >> 16) & 0xffu) << 8) | (b >> 24);
                                                len = 4u;
                                        } else {
                                                if (avail < 5u) {
                                                        win |= static_cast<unsigned long long>(nw) << (avail * 8u);
                                                        avail += 4u;
                                                        nw = src.next();
                                                }
                                                v   = uint32_t(win >> 8);
                                                len = 5u;
                                        }
                                        win >>= len * 8u;
                                        avail -= len;
                                        if (inDelta) {
                                                doc += v;
                                                rd.put(i, doc);
                                                sumd += doc;
                                        } else {
                                                rf.put(i, v);
                                                sumf += v;
                                        }
                                        ++c;
                                }
                        }
                        // the block's last doc is implied by the header/directory
                        rd.put(nd, cur.last);
                        sumd += cur.last;
                        rd.finish(cur.n);
                        rf.finish(cur.n);
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
                bufSel ^= 1u;
        }
}

cudaError_t launch_decode_google(const DevIndex &ix, const uint32_t *term_ids, const uint32_t *unit_base, const uint64_t *out_base, uint32_t nterms,
                                 uint32_t total_units, uint32_t *docids, uint32_t *freqs, unsigned long long *sums, int grid, cudaStream_t stream) {
        const size_t smem = size_t(kWarps) * 2 * kDecBufBytes;
        static const bool lockstep = getenv("TRN_DECODE_LOCKSTEP") && atoi(getenv("TRN_DECODE_LOCKSTEP")) != 0;
        if (lockstep)
                k_decode_google<true><<<grid, kThreads, smem, stream>>>(ix, term_ids, unit_base, out_base, nterms, total_units, docids, freqs, sums);
        else
                k_decode_google<false><<<grid, kThreads, smem, stream>>>(ix, term_ids, unit_base, out_base, nterms, total_units, docids, freqs, sums);
        return cudaGetLastError();
}
