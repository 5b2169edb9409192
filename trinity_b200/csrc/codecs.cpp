// Host-side encoders for the GOOGLE and LUCENE(FastPFor<4>) postings layouts + the load-time block directory.
// The two encoders restate the reference write path statement by statement (google_codec.cpp:9-176, lucene_codec.cpp:163-388;
// best_b follows fastpfor.h:143-171) onto std::vector — a byte-exact writer leaves little freedom, and the kernels must read exactly
// these bytes.  They are index-BUILD tooling (tests, the synthetic workload), not part of the query hot path; byte-exactness is
// pinned against the reference encoders by tests/test_codecs_cpu.py.  The block-directory half of the file is this repo's own.
#include "codecs.h"
#include "varbyte.h"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>

namespace trn {
namespace Codecs {

        // ======================================================================================= GOOGLE
        namespace Google {
                void Encoder::begin_term() {
                        auto &out               = sess->indexOut;
                        curBlockSize            = 0;
                        lastCommitedDocID       = 0;
                        prevBlockLastDocumentID = 0;
                        hitsData.clear();
                        skipListData.clear();
                        termDocuments = 0;
                        curTermOffset = uint32_t(out.size());
                        put_u16(out, 0); // skiplist entry count, patched in end_term()
                }

                void Encoder::begin_document(uint32_t documentID) {
                        if (documentID == 0 || documentID <= lastCommitedDocID)
                                throw std::invalid_argument("google encoder: document IDs must be > 0 and strictly ascending");
                        curDocID                 = documentID;
                        lastPos                  = 0;
                        curPayloadSize           = 0;
                        blockFreqs[curBlockSize] = 0;
                }

                void Encoder::new_hit(uint32_t pos, const uint8_t *payload, uint8_t payloadSize) {
                        if (!pos && !payloadSize)
                                return; // valid: a document with no positional hit (freq stays 0)
                        if (pos < lastPos || payloadSize > 8)
                                throw std::invalid_argument("google encoder: positions must be non-decreasing, payload <= 8 bytes");
                        const uint32_t delta = pos - lastPos;
                        ++blockFreqs[curBlockSize];
                        // TRACK_PAYLOADS layout: varbyte((delta<<1)|sizeChanged) [u8 newSize] payload
                        if (payloadSize != curPayloadSize) {
                                varbyte_put(hitsData, (delta << 1) | 1u);
                                hitsData.push_back(payloadSize);
                                curPayloadSize = payloadSize;
                        } else
                                varbyte_put(hitsData, delta << 1);
                        if (payloadSize)
                                hitsData.insert(hitsData.end(), payload, payload + payloadSize);
                        lastPos = pos;
                }

                void Encoder::end_document() {
                        docDeltas[curBlockSize++] = curDocID - lastCommitedDocID;
                        if (curBlockSize == blockDocs)
                                commit_block();
                        lastCommitedDocID = curDocID;
                        ++termDocuments;
                }

                void Encoder::commit_block() {
                        auto &         out   = sess->indexOut;
                        const uint32_t delta = curDocID - prevBlockLastDocumentID;

                        block.clear();
                        for (uint32_t i = 0; i + 1 < curBlockSize; ++i) // the last doc is implied by the header
                                varbyte_put(block, docDeltas[i]);
                        for (uint32_t i = 0; i < curBlockSize; ++i)
                                varbyte_put(block, blockFreqs[i]);

                        const uint32_t blockLength = uint32_t(block.size() + hitsData.size());

                        if (--skiplistEntryCountdown == 0) {
                                if (skipListData.size() / 8 < UINT16_MAX) {
                                        put_u32(skipListData, prevBlockLastDocumentID);
                                        put_u32(skipListData, uint32_t(out.size()) - curTermOffset);
                                }
                                skiplistEntryCountdown = skiplistStep;
                        }

                        varbyte_put(out, delta);
                        varbyte_put(out, blockLength);
                        out.push_back(uint8_t(curBlockSize));
                        out.insert(out.end(), block.begin(), block.end());
                        out.insert(out.end(), hitsData.begin(), hitsData.end());
                        hitsData.clear();

                        prevBlockLastDocumentID = curDocID;
                        curBlockSize            = 0;
                }

                void Encoder::end_term(term_index_ctx *tctx) {
                        auto &out = sess->indexOut;
                        if (curBlockSize)
                                commit_block();
                        const uint16_t entries = uint16_t(skipListData.size() / 8);
                        out.insert(out.end(), skipListData.begin(), skipListData.end());
                        out[curTermOffset]     = uint8_t(entries);
                        out[curTermOffset + 1] = uint8_t(entries >> 8);
                        tctx->offset           = curTermOffset;
                        tctx->size             = uint32_t(out.size()) - curTermOffset;
                        tctx->documents        = termDocuments;
                        skipListData.clear();
                }
        } // namespace Google

        // ======================================================================================= LUCENE
        namespace Lucene {
                namespace {
                        inline uint32_t bits_of(uint32_t v) {
                                return v ? 32 - uint32_t(__builtin_clz(v)) : 0;
                        }

                        // LSB-first packing of 32 values at `bit` bits into `bit` words
                        void pack32(const uint32_t *in, uint32_t *out, uint32_t bit, bool mask) {
                                for (uint32_t i = 0; i < bit; ++i)
                                        out[i] = 0;
                                if (!bit)
                                        return;
                                const uint32_t m = bit == 32 ? 0xffffffffu : ((1u << bit) - 1);
                                for (uint32_t i = 0; i < 32; ++i) {
                                        const uint32_t v  = mask ? (in[i] & m) : in[i];
                                        const uint32_t bp = i * bit, w = bp >> 5, sh = bp & 31;
                                        out[w] |= v << sh;
                                        if (sh + bit > 32)
                                                out[w + 1] |= v >> (32 - sh);
                                }
                        }

                        // FastPFor<4>::getBestBFromData cost model (fastpfor.h:143-171)
                        void best_b(const uint32_t *in, uint8_t &bestb, uint8_t &bestcexcept, uint8_t &maxb) {
                                uint32_t freqs[33] = {0};
                                for (uint32_t k = 0; k < BLOCK_SIZE; ++k)
                                        freqs[bits_of(in[k])]++;
                                uint32_t b = 32;
                                while (freqs[b] == 0)
                                        b--;
                                bestb             = uint8_t(b);
                                maxb              = uint8_t(b);
                                uint32_t bestcost = b * BLOCK_SIZE;
                                uint32_t cexcept  = 0;
                                bestcexcept       = 0;
                                for (int32_t bb = int32_t(b) - 1; bb >= 0; --bb) {
                                        cexcept += freqs[bb + 1];
                                        uint32_t thiscost = cexcept * 8 /*overheadofeachexcept*/ + cexcept * (maxb - bb) + uint32_t(bb) * BLOCK_SIZE + 8;
                                        if (maxb - bb == 1)
                                                thiscost -= cexcept;
                                        if (thiscost < bestcost) {
                                                bestcost    = thiscost;
                                                bestb       = uint8_t(bb);
                                                bestcexcept = uint8_t(cexcept);
                                        }
                                }
                        }
                } // namespace

                // FastPFor<4>::encodeArray of exactly one 128-value block (one "page" per call, lucene_codec.cpp:57-64)
                static uint32_t pfor_encode_page(const uint32_t *in, uint32_t *out) {
                        uint32_t *const initout = out;
                        *out++                  = BLOCK_SIZE; // length word (fastpfor.h:107)
                        uint32_t *const headerout = out++;    // wheremeta
                        uint8_t         bytes[4 + BLOCK_SIZE];
                        uint32_t        nb{0};
                        uint8_t         b, cexcept, maxb;
                        uint32_t        exceptions[BLOCK_SIZE];
                        uint32_t        nexc{0};

                        best_b(in, b, cexcept, maxb);
                        bytes[nb++] = b;
                        bytes[nb++] = cexcept;
                        if (cexcept > 0) {
                                bytes[nb++]            = maxb;
                                const uint32_t maxval = uint32_t(1ull << b);
                                for (uint32_t k = 0; k < BLOCK_SIZE; ++k) {
                                        if (in[k] >= maxval) {
                                                exceptions[nexc++] = in[k] >> b;
                                                bytes[nb++]        = uint8_t(k);
                                        }
                                }
                        }
                        for (uint32_t j = 0; j < BLOCK_SIZE; j += 32) {
                                pack32(in + j, out, b, true);
                                out += b;
                        }
                        headerout[0] = uint32_t(out - headerout);
                        *out++       = nb;
                        std::memset(out, 0, ((nb + 3) / 4) * 4); // deterministic padding (the reference leaves stale bytes; see note in tests)
                        std::memcpy(out, bytes, nb);
                        out += (nb + 3) / 4;

                        const uint32_t k      = uint32_t(maxb) - b; // exception width; k==1 carries no stream
                        uint32_t       bitmap = 0;
                        if (nexc && k >= 2)
                                bitmap |= 1u << (k - 1);
                        *out++ = bitmap;
                        if (bitmap) {
                                uint32_t padded[BLOCK_SIZE + 32] = {0};
                                std::memcpy(padded, exceptions, nexc * sizeof(uint32_t));
                                *out++ = nexc;
                                uint32_t j{0};
                                for (; j < nexc; j += 32) {
                                        pack32(padded + j, out, k, false);
                                        out += k;
                                }
                                out -= (j - nexc) * k / 32;
                        }
                        return uint32_t(out - initout);
                }

                void ints_encode(const uint32_t *values, std::vector<uint8_t> &out) {
                        bool eq{true};
                        for (uint32_t i = 1; i < BLOCK_SIZE; ++i)
                                if (values[i] != values[0]) {
                                        eq = false;
                                        break;
                                }
                        if (eq) {
                                out.push_back(0);
                                varbyte_put(out, values[0]);
                                return;
                        }
                        uint32_t       page[2 * BLOCK_SIZE + 64];
                        const uint32_t l = pfor_encode_page(values, page);
                        out.push_back(uint8_t(l));
                        const auto *bytes = reinterpret_cast<const uint8_t *>(page);
                        out.insert(out.end(), bytes, bytes + size_t(l) * 4);
                }

                void Encoder::begin_term() {
                        lastDocID = totalHits = sumHits = buffered = termDocuments = 0;
                        termIndexOffset        = uint32_t(sess->indexOut.size());
                        termPositionsOffset    = uint32_t(sess->positionsOut.size());
                        lastHitsBlockOffset    = 0;
                        lastHitsBlockTotalHits = 0;
                        skiplistCountdown      = SKIPLIST_STEP;
                        skiplist.clear();
                        payloadsBuf.clear();
                        // chunk header: u32 hitsDataOffset, u32 sumHits, u32 positionsChunkSize, u16 skiplistSize (patched in end_term)
                        put_u32(sess->indexOut, termPositionsOffset);
                        put_u32(sess->indexOut, 0);
                        put_u32(sess->indexOut, 0);
                        put_u16(sess->indexOut, 0);
                }

                void Encoder::output_block() {
                        if (--skiplistCountdown == 0) {
                                if (skiplist.size() < UINT16_MAX)
                                        skiplist.push_back(cur_block);
                                skiplistCountdown = SKIPLIST_STEP;
                        }
                        ints_encode(docDeltas, sess->indexOut);
                        ints_encode(docFreqs, sess->indexOut);
                        buffered = 0;
                }

                void Encoder::begin_document(uint32_t documentID) {
                        if (documentID <= lastDocID)
                                throw std::invalid_argument("lucene encoder: document IDs must be > 0 and strictly ascending");
                        if (buffered == BLOCK_SIZE)
                                output_block();
                        if (!buffered) {
                                cur_block.indexOffset            = uint32_t(sess->indexOut.size()) - termIndexOffset;
                                cur_block.lastDocID              = lastDocID;
                                cur_block.totalDocumentsSoFar    = termDocuments;
                                cur_block.lastHitsBlockOffset    = lastHitsBlockOffset;
                                cur_block.lastHitsBlockTotalHits = lastHitsBlockTotalHits;
                                cur_block.curHitsBlockHits       = uint16_t(totalHits);
                        }
                        docDeltas[buffered] = documentID - lastDocID;
                        docFreqs[buffered]  = 0;
                        ++termDocuments;
                        lastDocID    = documentID;
                        lastPosition = 0;
                }

                void Encoder::new_hit(uint32_t pos, const uint8_t *payload, uint8_t payloadSize) {
                        if (!pos && !payloadSize)
                                return;
                        if (pos < lastPosition || payloadSize > 8)
                                throw std::invalid_argument("lucene encoder: positions must be non-decreasing, payload <= 8 bytes");
                        ++docFreqs[buffered];
                        hitPosDeltas[totalHits]    = pos - lastPosition;
                        hitPayloadSizes[totalHits] = payloadSize;
                        lastPosition               = pos;
                        if (payloadSize)
                                payloadsBuf.insert(payloadsBuf.end(), payload, payload + payloadSize);
                        if (++totalHits == BLOCK_SIZE) {
                                auto &po = sess->positionsOut;
                                sumHits += totalHits;
                                ints_encode(hitPosDeltas, po);
                                ints_encode(hitPayloadSizes, po);
                                varbyte_put(po, uint32_t(payloadsBuf.size()));
                                po.insert(po.end(), payloadsBuf.begin(), payloadsBuf.end());
                                payloadsBuf.clear();
                                lastHitsBlockTotalHits = sumHits;
                                lastHitsBlockOffset    = uint32_t(po.size()) - termPositionsOffset;
                                totalHits              = 0;
                        }
                }

                void Encoder::end_document() {
                        ++buffered;
                }

                void Encoder::end_term(term_index_ctx *out) {
                        auto &io = sess->indexOut;
                        auto &po = sess->positionsOut;
                        sumHits += totalHits;
                        if (buffered == BLOCK_SIZE)
                                output_block();
                        else {
                                for (uint32_t i = 0; i < buffered; ++i) {
                                        varbyte_put(io, docDeltas[i]);
                                        varbyte_put(io, docFreqs[i]);
                                }
                        }
                        if (totalHits) {
                                uint8_t lastPayloadLen{0};
                                for (uint32_t i = 0; i < totalHits; ++i) {
                                        const uint8_t pl = uint8_t(hitPayloadSizes[i]);
                                        if (pl != lastPayloadLen) {
                                                lastPayloadLen = pl;
                                                varbyte_put(po, (hitPosDeltas[i] << 1) | 1u);
                                                po.push_back(pl);
                                        } else
                                                varbyte_put(po, hitPosDeltas[i] << 1);
                                }
                                po.insert(po.end(), payloadsBuf.begin(), payloadsBuf.end());
                                payloadsBuf.clear();
                        }
                        auto patch32 = [&](uint32_t at, uint32_t v) {
                                io[at]     = uint8_t(v);
                                io[at + 1] = uint8_t(v >> 8);
                                io[at + 2] = uint8_t(v >> 16);
                                io[at + 3] = uint8_t(v >> 24);
                        };
                        const uint16_t skiplistSize = uint16_t(skiplist.size());
                        patch32(termIndexOffset + 4, sumHits);
                        patch32(termIndexOffset + 8, uint32_t(po.size()) - termPositionsOffset);
                        io[termIndexOffset + 12] = uint8_t(skiplistSize);
                        io[termIndexOffset + 13] = uint8_t(skiplistSize >> 8);
                        for (const auto &e : skiplist) {
                                put_u32(io, e.indexOffset);
                                put_u32(io, e.lastDocID);
                                put_u32(io, e.lastHitsBlockOffset);
                                put_u32(io, e.totalDocumentsSoFar);
                                put_u32(io, e.lastHitsBlockTotalHits);
                                put_u16(io, e.curHitsBlockHits);
                        }
                        skiplist.clear();
                        out->documents = termDocuments;
                        out->offset    = termIndexOffset;
                        out->size      = uint32_t(io.size()) - termIndexOffset;
                }
        } // namespace Lucene

        Encoder *new_encoder(IndexSession *s) {
                if (s->codec == Codec::Google)
                        return new Google::Encoder(s);
                return new Lucene::Encoder(s);
        }
} // namespace Codecs

// =========================================================================================== block directory
namespace {
        // varbyte read that never leaves [p, end): the directory walks bytes that come from files
        inline uint32_t vb_checked(const uint8_t *&p, const uint8_t *end, const char *what) {
                if (p >= end || p + varbyte_len(*p) > end)
                        throw std::runtime_error(std::string(what) + ": varbyte code past the end of its chunk");
                return varbyte_get(p);
        }

        // host decode of one Lucene int-block (sum only is needed for the directory; full values for tests)
        const uint8_t *lucene_ints_decode(const uint8_t *p, const uint8_t *end, uint32_t *values) {
                using Codecs::Lucene::BLOCK_SIZE;
                if (p >= end)
                        throw std::runtime_error("lucene: int-block past chunk end");
                const uint32_t L = *p++;
                if (L == 0) {
                        const uint32_t v = vb_checked(p, end, "lucene");
                        for (uint32_t i = 0; i < BLOCK_SIZE; ++i)
                                values[i] = v;
                        return p;
                }
                if (p + size_t(L) * 4 > end)
                        throw std::runtime_error("lucene: PFor page past chunk end");
                auto w = [&](uint32_t i) {
                        if (i >= L)
                                throw std::runtime_error("lucene: PFor page refers to a word outside itself");
                        return get_u32(p + size_t(i) * 4);
                };
                if (w(0) != BLOCK_SIZE)
                        throw std::runtime_error("lucene: PFor page length word != 128");
                const uint32_t wheremeta = w(1);
                const uint32_t b         = (wheremeta - 1) / 4;
                if (wheremeta != 1 + 4 * b || b > 32 || 2 + wheremeta > L)
                        throw std::runtime_error("lucene: malformed PFor page header");
                for (uint32_t i = 0; i < BLOCK_SIZE; ++i) {
                        if (!b) {
                                values[i] = 0;
                                continue;
                        }
                        const uint32_t g = i >> 5, j = i & 31, bp = j * b, wi = 2 + g * b + (bp >> 5), sh = bp & 31;
                        uint64_t       x = w(wi);
                        if (sh + b > 32)
                                x |= uint64_t(w(wi + 1)) << 32;
                        values[i] = uint32_t((x >> sh) & (b == 32 ? 0xffffffffull : ((1ull << b) - 1)));
                }
                const uint32_t meta     = 1 + wheremeta;
                const uint32_t bytesize = w(meta);
                const uint8_t *bytes    = p + size_t(meta + 1) * 4;
                const size_t   pageLeft = size_t(L) * 4 - size_t(meta + 1) * 4; // bytes of the page behind the bytesize word
                if (pageLeft < 2 || bytesize > pageLeft)
                        throw std::runtime_error("lucene: PFor byte container outside its page");
                if (bytes[0] != b)
                        throw std::runtime_error("lucene: PFor b mismatch");
                const uint32_t cexcept = bytes[1];
                if (cexcept) {
                        if (size_t(3) + cexcept > pageLeft)
                                throw std::runtime_error("lucene: PFor exception positions outside the page");
                        const uint32_t maxbits = bytes[2];
                        if (maxbits <= b || maxbits > 32)
                                throw std::runtime_error("lucene: PFor exception width out of range");
                        const uint32_t k       = maxbits - b;
                        const uint32_t excw    = meta + 1 + (bytesize + 3) / 4; // bitmap word
                        for (uint32_t e = 0; e < cexcept; ++e) {
                                const uint32_t pos = bytes[3 + e];
                                if (pos >= BLOCK_SIZE)
                                        throw std::runtime_error("lucene: PFor exception position >= 128");
                                uint32_t       ev{1};
                                if (k > 1) {
                                        const uint32_t base = excw + 2; // after bitmap + count
                                        const uint32_t bp = e * k, wi = base + (bp >> 5), sh = bp & 31;
                                        uint64_t       x = w(wi);
                                        if (sh + k > 32)
                                                x |= uint64_t(w(wi + 1)) << 32;
                                        ev = uint32_t((x >> sh) & (k == 32 ? 0xffffffffull : ((1ull << k) - 1)));
                                }
                                values[pos] |= ev << b;
                        }
                }
                return p + size_t(L) * 4;
        }

        // blockN: documents per full block of this term (0 when the term has a single block: it cannot tell)
        void dir_google_term(const uint8_t *index, const term_index_ctx &t, std::vector<uint32_t> &last, std::vector<uint32_t> &off, uint32_t &firstDoc, uint32_t &blockN) {
                const uint32_t N = Codecs::Google::MAX_N;
                firstDoc = 0;
                blockN   = 0;
                if (!t.size) {
                        if (t.documents)
                                throw std::runtime_error("google: term with documents but empty chunk");
                        return;
                }
                const uint8_t *base     = index + t.offset;
                if (t.size < 2)
                        throw std::runtime_error("google: chunk shorter than its header");
                const uint32_t entries  = get_u16(base);
                if (size_t(entries) * 8 + 2 > t.size)
                        throw std::runtime_error("google: skiplist larger than the chunk");
                const uint8_t *chunkEnd = base + t.size - size_t(entries) * 8;
                const uint8_t *p        = base + 2;
                uint32_t       prev{0}, docs{0};
                while (p < chunkEnd) {
                        const uint32_t delta = vb_checked(p, chunkEnd, "google");
                        const uint32_t blen  = vb_checked(p, chunkEnd, "google");
                        if (p >= chunkEnd)
                                throw std::runtime_error("google: block header past the end of its chunk");
                        const uint32_t n     = *p++;
                        if (n == 0 || n > N)
                                throw std::runtime_error("google: bad block doc count");
                        if (blen > size_t(chunkEnd - p))
                                throw std::runtime_error("google: block longer than its chunk");
                        if (last.empty()) {
                                const uint8_t *q = p;
                                firstDoc         = n > 1 ? vb_checked(q, chunkEnd, "google") : delta;
                        }
                        prev += delta;
                        last.push_back(prev);
                        off.push_back(uint32_t(p - index));
                        docs += n;
                        if (docs != t.documents) { // not the term's last block: all of these must have the same size
                                if (!blockN)
                                        blockN = n;
                                else if (n != blockN)
                                        throw std::runtime_error("google: non-final block is not full (unsupported by the GPU directory)");
                        } else if (blockN && n > blockN)
                                throw std::runtime_error("google: final block larger than the term's block size");
                        p += blen;
                }
                if (p != chunkEnd || docs != t.documents)
                        throw std::runtime_error("google: chunk walk did not end at chunkEnd / documents mismatch");
                last.push_back(UINT32_MAX);
                off.push_back(uint32_t(chunkEnd - index));
        }

        void dir_lucene_term(const uint8_t *index, const term_index_ctx &t, std::vector<uint32_t> &last, std::vector<uint32_t> &off, uint32_t &firstDoc) {
                using Codecs::Lucene::BLOCK_SIZE;
                firstDoc = 0;
                if (!t.size) {
                        if (t.documents)
                                throw std::runtime_error("lucene: term with documents but empty chunk");
                        return;
                }
                const uint8_t *base = index + t.offset;
                if (t.size < 14)
                        throw std::runtime_error("lucene: chunk shorter than its header");
                const uint32_t skipn    = get_u16(base + 12);
                if (size_t(skipn) * 22 + 14 > t.size)
                        throw std::runtime_error("lucene: skiplist larger than the chunk");
                const uint8_t *chunkEnd = base + t.size - size_t(skipn) * 22;
                // length-byte hop over one int-block, bounds-checked
                auto hop = [&](const uint8_t *&q) {
                        if (q >= chunkEnd)
                                throw std::runtime_error("lucene: int-block past chunk end");
                        const uint32_t L = *q++;
                        if (L == 0)
                                (void)vb_checked(q, chunkEnd, "lucene");
                        else {
                                if (size_t(L) * 4 > size_t(chunkEnd - q))
                                        throw std::runtime_error("lucene: PFor page past chunk end");
                                q += size_t(L) * 4;
                        }
                };
                const uint8_t *skip     = chunkEnd;
                const uint8_t *p        = base + 14;
                const uint32_t nfull    = t.documents / BLOCK_SIZE;
                const uint32_t tail     = t.documents % BLOCK_SIZE;
                uint32_t       prev{0};
                uint32_t       vals[BLOCK_SIZE];
                for (uint32_t blk = 0; blk < nfull; ++blk) {
                        off.push_back(uint32_t(p - index));
                        const bool haveNext = blk + 1 < skipn; // skiplist entry blk+1 holds this block's last docID
                        if (blk < skipn) {
                                // cross-check the on-disk skiplist entry for this block
                                if (get_u32(skip + size_t(blk) * 22) != uint32_t(p - base) || get_u32(skip + size_t(blk) * 22 + 4) != prev)
                                        throw std::runtime_error("lucene: skiplist entry disagrees with block walk");
                        }
                        if (haveNext && blk != 0) {
                                // fast path: skip the two int-blocks by their length bytes
                                hop(p);
                                hop(p);
                                prev = get_u32(skip + size_t(blk + 1) * 22 + 4);
                        } else {
                                p = lucene_ints_decode(p, chunkEnd, vals);
                                if (blk == 0)
                                        firstDoc = vals[0];
                                uint32_t s{0};
                                for (uint32_t i = 0; i < BLOCK_SIZE; ++i)
                                        s += vals[i];
                                prev += s;
                                hop(p); // the freqs block
                        }
                        last.push_back(prev);
                }
                if (tail) {
                        off.push_back(uint32_t(p - index));
                        for (uint32_t i = 0; i < tail; ++i) {
                                const uint32_t d = vb_checked(p, chunkEnd, "lucene");
                                (void)vb_checked(p, chunkEnd, "lucene");
                                prev += d;
                                if (nfull == 0 && i == 0)
                                        firstDoc = d;
                        }
                        last.push_back(prev);
                }
                if (p != chunkEnd)
                        throw std::runtime_error("lucene: chunk walk did not end at the skiplist");
                last.push_back(UINT32_MAX);
                off.push_back(uint32_t(chunkEnd - index));
        }
} // namespace

// exposed for tests (host decode of a Lucene int-block)
const uint8_t *lucene_ints_decode_host(const uint8_t *p, const uint8_t *end, uint32_t *values) {
        return lucene_ints_decode(p, end, values);
}

void build_block_directory(Codec codec, const uint8_t *index, uint64_t nbytes, const term_index_ctx *terms, uint32_t nterms, int threads, BlockDirectory &out) {
        struct PerTerm {
                std::vector<uint32_t> last, off;
                uint32_t              firstDoc{0}, blockN{0};
        };
        std::vector<PerTerm>  per(nterms);
        std::atomic<uint32_t> next{0};
        std::string           err;
        std::atomic<bool>     failed{false};
        auto                  worker = [&] {
                for (;;) {
                        const uint32_t i = next.fetch_add(1);
                        if (i >= nterms || failed.load())
                                break;
                        try {
                                if (uint64_t(terms[i].offset) + terms[i].size > nbytes)
                                        throw std::runtime_error("term chunk exceeds index size");
                                if (codec == Codec::Google)
                                        dir_google_term(index, terms[i], per[i].last, per[i].off, per[i].firstDoc, per[i].blockN);
                                else
                                        dir_lucene_term(index, terms[i], per[i].last, per[i].off, per[i].firstDoc);
                        } catch (const std::exception &e) {
                                bool exp{false};
                                if (failed.compare_exchange_strong(exp, true))
                                        err = std::string("term ") + std::to_string(i) + ": " + e.what();
                        }
                }
        };
        if (threads < 1)
                threads = 1;
        std::vector<std::thread> ths;
        for (int i = 1; i < threads; ++i)
                ths.emplace_back(worker);
        worker();
        for (auto &t : ths)
                t.join();
        if (failed.load())
                throw std::runtime_error("build_block_directory: " + err);

        // one block size per index (GOOGLE: google_codec.h:18 says 32; a multi-block term shows it)
        out.block_docs = codec == Codec::Lucene ? Codecs::Lucene::BLOCK_SIZE : 0u;
        if (codec == Codec::Google) {
                for (auto &p : per)
                        if (p.blockN) {
                                if (out.block_docs && out.block_docs != p.blockN)
                                        throw std::runtime_error("build_block_directory: terms disagree on the block size");
                                out.block_docs = p.blockN;
                        }
                if (!out.block_docs)
                        out.block_docs = Codecs::Google::N;
        }
        size_t total{0};
        for (auto &p : per)
                total += p.last.size();
        if (total >= (1ull << 32))
                throw std::runtime_error("block directory exceeds 2^32 entries");
        out.blk_last.resize(total);
        out.blk_off.resize(total);
        out.terms.resize(nterms);
        size_t at{0};
        for (uint32_t i = 0; i < nterms; ++i) {
                auto &p  = per[i];
                auto &td = out.terms[i];
                td.documents = terms[i].documents;
                td.dir_begin = uint32_t(at);
                td.nblocks   = p.last.empty() ? 0 : uint32_t(p.last.size() - 1);
                td.first_doc = p.firstDoc;
                td.last_doc  = td.nblocks ? p.last[td.nblocks - 1] : 0;
                std::copy(p.last.begin(), p.last.end(), out.blk_last.begin() + at);
                std::copy(p.off.begin(), p.off.end(), out.blk_off.begin() + at);
                at += p.last.size();
                std::vector<uint32_t>().swap(p.last);
                std::vector<uint32_t>().swap(p.off);
        }
        // sparse docID -> block tables (see codecs.h): sized first, then filled in parallel
        size_t tfTotal{0};
        for (uint32_t i = 0; i < nterms; ++i) {
                auto &td = out.terms[i];
                td.tf_begin = uint32_t(tfTotal);
                td.tf_base = td.tf_n = 0;
                td.tf_shift          = kDirNoTable;
                if (td.nblocks <= kDirNoTableBlocks)
                        continue;
                uint32_t s = kDirMinShift;
                while (s < 31 && (uint64_t(td.last_doc >> s) - (td.first_doc >> s) + 1) * kDirBlocksPerEntry > td.nblocks)
                        ++s;
                td.tf_shift = s;
                td.tf_base  = td.first_doc >> s;
                td.tf_n     = (td.last_doc >> s) - td.tf_base + 1;
                tfTotal += size_t(td.tf_n) + 1;
                if (tfTotal >= (1ull << 32))
                        throw std::runtime_error("block directory: tile table exceeds 2^32 entries");
        }
        out.tile_first.resize(tfTotal);
        {
                std::atomic<uint32_t> nx{0};
                auto                  fill = [&] {
                        for (;;) {
                                const uint32_t i = nx.fetch_add(1);
                                if (i >= nterms)
                                        break;
                                const auto &td = out.terms[i];
                                if (td.tf_shift == kDirNoTable)
                                        continue;
                                const uint32_t *bl = out.blk_last.data() + td.dir_begin;
                                uint32_t *      o  = out.tile_first.data() + td.tf_begin;
                                uint32_t        b{0};
                                for (uint32_t j = 0; j <= td.tf_n; ++j) {
                                        const uint64_t lo = uint64_t(td.tf_base + j) << td.tf_shift;
                                        while (b < td.nblocks && bl[b] < lo)
                                                ++b;
                                        o[j] = b;
                                }
                        }
                };
                std::vector<std::thread> fths;
                for (int i = 1; i < threads; ++i)
                        fths.emplace_back(fill);
                fill();
                for (auto &t : fths)
                        t.join();
        }
}

} // namespace trn

namespace trn {

void build_hits_directory(const uint8_t *index, uint64_t nbytes, const uint8_t *hits, uint64_t hbytes, const term_index_ctx *terms, uint32_t nterms,
                          const BlockDirectory &dir, int threads, HitsDirectory &out) {
        using Codecs::Lucene::BLOCK_SIZE;
        if (dir.terms.size() != nterms)
                throw std::runtime_error("hits directory: block directory of another index");
        out.hit_base.assign(dir.blk_last.size(), 0);
        out.hb_begin.assign(nterms, 0);
        out.sum_hits.assign(nterms, 0);
        // entries of hblk_off per term are known from the chunk headers
        uint64_t total{0};
        for (uint32_t i = 0; i < nterms; ++i) {
                const auto &t = terms[i];
                out.hb_begin[i] = uint32_t(total);
                if (!t.documents || !t.size)
                        continue;
                if (uint64_t(t.offset) + t.size > nbytes || t.size < 14)
                        throw std::runtime_error("lucene: term chunk outside the index");
                out.sum_hits[i] = get_u32(index + t.offset + 4);
                total += uint64_t(out.sum_hits[i] / BLOCK_SIZE) + 2;
                if (total >= (1ull << 32))
                        throw std::runtime_error("hits directory: more than 2^32 hit blocks");
        }
        out.hblk_off.assign(total, 0);
        std::atomic<uint32_t> next{0};
        std::atomic<bool>     failed{false};
        std::string           err;
        std::mutex            mu;
        auto                  worker = [&] {
                uint32_t vals[BLOCK_SIZE];
                for (;;) {
                        const uint32_t i = next.fetch_add(1);
                        if (i >= nterms || failed.load())
                                break;
                        const auto &t = terms[i];
                        if (!t.documents || !t.size)
                                continue;
                        try {
                                const auto &   td       = dir.terms[i];
                                const uint8_t *base     = index + t.offset;
                                const uint32_t hitsOff  = get_u32(base), sumHits = out.sum_hits[i], skipn = uint32_t(base[12]) | (uint32_t(base[13]) << 8);
                                const uint8_t *chunkEnd = base + t.size - size_t(skipn) * 22;
                                // ---- hits before every document block: the freqs of the block walk
                                const uint32_t nfull = t.documents / BLOCK_SIZE, tail = t.documents % BLOCK_SIZE;
                                if (td.nblocks != nfull + (tail ? 1u : 0u))
                                        throw std::runtime_error("hits directory: block count disagrees with the block directory");
                                uint64_t run{0};
                                for (uint32_t b = 0; b < nfull; ++b) {
                                        const uint8_t *p = index + dir.blk_off[td.dir_begin + b];
                                        out.hit_base[td.dir_begin + b] = uint32_t(run);
                                        // the deltas int-block is skipped by its length byte, the freqs int-block decoded
                                        if (p >= chunkEnd)
                                                throw std::runtime_error("lucene: int-block past chunk end");
                                        const uint32_t L = *p++;
                                        if (L == 0)
                                                (void)vb_checked(p, chunkEnd, "lucene");
                                        else
                                                p += size_t(L) * 4;
                                        (void)lucene_ints_decode(p, chunkEnd, vals);
                                        for (uint32_t k = 0; k < BLOCK_SIZE; ++k)
                                                run += vals[k];
                                }
                                if (tail) {
                                        const uint8_t *p = index + dir.blk_off[td.dir_begin + nfull];
                                        out.hit_base[td.dir_begin + nfull] = uint32_t(run);
                                        for (uint32_t k = 0; k < tail; ++k) {
                                                (void)vb_checked(p, chunkEnd, "lucene");
                                                run += vb_checked(p, chunkEnd, "lucene");
                                        }
                                }
                                out.hit_base[td.dir_begin + td.nblocks] = uint32_t(run); // sentinel
                                if (run != sumHits)
                                        throw std::runtime_error("lucene: the freqs of a term do not add up to its sumHits");
                                // ---- the term's hit blocks
                                if (!sumHits)
                                        continue;
                                if (uint64_t(hitsOff) >= hbytes)
                                        throw std::runtime_error("lucene: hits offset outside hits.data");
                                const uint8_t *hend = hits + hbytes, *p = hits + hitsOff;
                                uint32_t *     o    = out.hblk_off.data() + out.hb_begin[i];
                                auto           hop  = [&](const uint8_t *&q) {
                                        if (q >= hend)
                                                throw std::runtime_error("lucene: hit block past the end of hits.data");
                                        const uint32_t L = *q++;
                                        if (L == 0)
                                                (void)vb_checked(q, hend, "lucene hits");
                                        else {
                                                if (size_t(L) * 4 > size_t(hend - q))
                                                        throw std::runtime_error("lucene: hit block past the end of hits.data");
                                                q += size_t(L) * 4;
                                        }
                                };
                                const uint32_t nfh = sumHits / BLOCK_SIZE;
                                for (uint32_t h = 0; h < nfh; ++h) {
                                        o[h] = uint32_t(p - hits);
                                        hop(p); // position deltas
                                        hop(p); // payload sizes
                                        const uint32_t plen = vb_checked(p, hend, "lucene hits");
                                        if (plen > size_t(hend - p))
                                                throw std::runtime_error("lucene: payloads past the end of hits.data");
                                        p += plen;
                                }
                                o[nfh] = uint32_t(p - hits);
                                for (uint32_t k = 0; k < sumHits % BLOCK_SIZE; ++k) {
                                        const uint32_t v = vb_checked(p, hend, "lucene hits");
                                        if (v & 1u) {
                                                if (p >= hend)
                                                        throw std::runtime_error("lucene: hit tail past the end of hits.data");
                                                ++p;
                                        }
                                }
                                o[nfh + 1] = uint32_t(p - hits);
                        } catch (const std::exception &e) {
                                std::lock_guard<std::mutex> g(mu);
                                if (!failed.exchange(true))
                                        err = std::string("term ") + std::to_string(i) + ": " + e.what();
                        }
                }
        };
        std::vector<std::thread> pool;
        const int                n = std::max(1, std::min(threads, 64));
        for (int k = 1; k < n; ++k)
                pool.emplace_back(worker);
        worker();
        for (auto &th : pool)
                th.join();
        if (failed.load())
                throw std::runtime_error(err);
}

} // namespace trn
