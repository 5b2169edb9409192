/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * Plain-C CPU restatement of the reference's hot path (postings decode -> docset algebra -> BM25), written from the
 * reference's behaviour, every function citing the reference file:line it follows.  It is PINNED against the reference
 * itself: tests/test_oracle_cpu.py checks every function here against oracle/_ref/libtrinity_ref.so (the reference's own
 * code compiled in place by oracle/build_ref.sh) and against the golden vectors in tests/golden/ that were generated from it.
 * The reference ships no tests or golden vectors of its own for this path, so "pinned" here means
 * "pinned by executing the reference".
 *
 * The docset algebra is restated the simplest possible way (one byte + one double per document) because its job is to be
 * obviously right, not fast: it is the checker, never the thing measured.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- prefix varbyte: Switch/switch_compiler_aux.h:53-81 (varbyte_get32) ---- */
static uint32_t vb_get(const uint8_t **pp) {
        const uint8_t *p = *pp;
        uint32_t       x = *p++;
        if (!(x & 0x80u)) {
        } else if (!(x & 0x40u)) {
                x = ((x & 0x3fu) << 8) | p[0];
                p += 1;
        } else if (!(x & 0x20u)) {
                x = ((x & 0x1fu) << 16) | p[0] | ((uint32_t)p[1] << 8);
                p += 2;
        } else if (!(x & 0x10u)) {
                x = ((x & 0x0fu) << 24) | ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2];
                p += 3;
        } else {
                x = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
                p += 4;
        }
        *pp = p;
        return x;
}
static uint32_t rd32(const uint8_t *p) {
        return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static uint32_t rd16(const uint8_t *p) {
        return p[0] | ((uint32_t)p[1] << 8);
}

/* ---- GOOGLE codec: Decoder::init (google_codec.cpp:936-983), unpack_block (:596-639), next (:777-819), skip_block_doc (:497-531).
 * Walks the chunk exactly like a sequence of next() calls, INCLUDING the varbyte walk over every document's hits that the
 * reference needs to find the next block (TRACK_PAYLOADS layout, google_codec.cpp:58-71).  Returns #postings or -1. */
int64_t orc_decode_google(const uint8_t *chunk, uint32_t len, uint32_t *docids, uint32_t *freqs, uint64_t cap) {
        if (len == 0)
                return 0;
        const uint32_t entries  = rd16(chunk);
        const uint8_t *end      = chunk + len - (size_t)entries * 8; /* chunkEnd = start of skiplist */
        const uint8_t *p        = chunk + 2;
        uint32_t       prevLast = 0;
        uint64_t       n        = 0;
        uint32_t       docs[32], fr[32];
        while (p < end) {
                const uint32_t last = prevLast + vb_get(&p);
                (void)vb_get(&p); /* block length: the reference's next() does not use it, it walks the hits */
                const uint32_t cnt = *p++;
                if (cnt == 0 || cnt > 32)
                        return -1;
                uint32_t id = prevLast;
                for (uint32_t i = 0; i + 1 < cnt; ++i) {
                        id += vb_get(&p);
                        docs[i] = id;
                }
                docs[cnt - 1] = last;
                for (uint32_t i = 0; i < cnt; ++i)
                        fr[i] = vb_get(&p);
                for (uint32_t i = 0; i < cnt; ++i) {
                        if (n < cap) {
                                docids[n] = docs[i];
                                freqs[n]  = fr[i] & 0xffffu; /* PostingsListIterator::freq is tokenpos_t = uint16_t (codecs.h:217, common.h:46) */
                        }
                        ++n;
                        uint8_t payloadSize = 0;
                        for (uint32_t h = 0; h < fr[i]; ++h) { /* skip_block_doc */
                                const uint32_t step = vb_get(&p);
                                if (step & 1u)
                                        payloadSize = *p++;
                                p += payloadSize;
                        }
                }
                prevLast = last;
        }
        return p == end ? (int64_t)n : -1;
}

/* ---- LUCENE codec int-block: ints_decode (lucene_codec.cpp:69-100) + FastPFor<4>::__decodeArray (fastpfor.h:222-270),
 * fastunpack (bitpackinghelpers.h:15-120: value i of a 32-value group sits at bits [i*b, (i+1)*b) LSB-first) and
 * packingvector<32>::unpackmetight (packingvectors.h:34-58) for the exception stream. */
static uint32_t bits_at(const uint8_t *words, uint32_t bitpos, uint32_t nbits) {
        if (nbits == 0)
                return 0;
        uint64_t       x  = rd32(words + (size_t)(bitpos >> 5) * 4);
        const uint32_t sh = bitpos & 31u;
        if (sh + nbits > 32)
                x |= (uint64_t)rd32(words + (size_t)(bitpos >> 5) * 4 + 4) << 32;
        x >>= sh;
        return nbits == 32 ? (uint32_t)x : (uint32_t)(x & ((1ull << nbits) - 1));
}

static const uint8_t *orc_ints_decode(const uint8_t *p, uint32_t *values) {
        const uint32_t L = *p++;
        if (L == 0) {
                const uint32_t v = vb_get(&p);
                for (int i = 0; i < 128; ++i)
                        values[i] = v;
                return p;
        }
        /* page words: [0]=nvalue(128) [1]=wheremeta [2..] 4 groups x b packed words ; meta at word 1+wheremeta */
        const uint32_t wheremeta = rd32(p + 4);
        const uint8_t *meta      = p + (size_t)(1 + wheremeta) * 4;
        const uint32_t bytesize  = rd32(meta);
        const uint8_t *bytep     = meta + 4;
        const uint8_t *inexcept  = bytep + (size_t)((bytesize + 3) / 4) * 4;
        const uint32_t bitmap    = rd32(inexcept);
        inexcept += 4;
        /* exception streams, one per set bit k-1 (k = maxbits - b in 2..32): u32 count, then count values at k bits, tight */
        const uint8_t *excStream[33] = {0};
        uint32_t       excIdx[33]    = {0};
        for (uint32_t k = 2; k <= 32; ++k) {
                if (bitmap & (1u << (k - 1))) {
                        const uint32_t cnt = rd32(inexcept);
                        excStream[k]       = inexcept + 4;
                        inexcept += 4 + (size_t)((cnt * k + 31) / 32) * 4;
                }
        }
        const uint32_t b       = *bytep++;
        const uint32_t cexcept = *bytep++;
        for (uint32_t g = 0; g < 4; ++g)
                for (uint32_t j = 0; j < 32; ++j)
                        values[g * 32 + j] = bits_at(p + 8 + (size_t)g * b * 4, j * b, b);
        if (cexcept) {
                const uint32_t maxbits = *bytep++;
                const uint32_t k       = maxbits - b;
                for (uint32_t e = 0; e < cexcept; ++e) {
                        const uint32_t pos = *bytep++;
                        if (k == 1)
                                values[pos] |= 1u << b;
                        else
                                values[pos] |= bits_at(excStream[k], (excIdx[k]++) * k, k) << b;
                }
        }
        return p + (size_t)L * 4;
}

/* Lucene::Decoder::init (lucene_codec.cpp:896-932) + refill_documents / next (:515-594): 14-byte header, documents/128 full
 * blocks (deltas int-block, freqs int-block), then (documents % 128) varbyte (delta, freq) pairs, skiplist at the end. */
int64_t orc_decode_lucene(const uint8_t *chunk, uint32_t len, uint32_t documents, uint32_t *docids, uint32_t *freqs, uint64_t cap) {
        if (len == 0)
                return 0;
        const uint32_t skipn = rd16(chunk + 12);
        const uint8_t *end   = chunk + len - (size_t)skipn * 22;
        const uint8_t *p     = chunk + 14;
        uint32_t       id = 0, d[128], f[128];
        uint64_t       n = 0;
        for (uint32_t blk = 0; blk < documents / 128; ++blk) {
                p = orc_ints_decode(p, d);
                p = orc_ints_decode(p, f);
                for (int i = 0; i < 128; ++i) {
                        id += d[i];
                        if (n < cap) {
                                docids[n] = id;
                                freqs[n]  = f[i] & 0xffffu;
                        }
                        ++n;
                }
        }
        for (uint32_t i = 0; i < documents % 128; ++i) {
                id += vb_get(&p);
                const uint32_t fr = vb_get(&p);
                if (n < cap) {
                        docids[n] = id;
                        freqs[n]  = fr & 0xffffu;
                }
                ++n;
        }
        return p == end ? (int64_t)n : -1;
}

/* ---- positions (restated ahead of the device path) ----
 * GOOGLE: Decoder::materialize_hits (google_codec.cpp:533-594): the hits of a block's documents follow its freqs; per hit
 * varbyte((posDelta << 1) | payloadSizeChanged) [u8 payloadSize] payload[payloadSize]; positions restart at 0 for every document.
 * positions[] receives freq(doc) entries per document in list order.  Returns the number of positions or -1. */
int64_t orc_positions_google(const uint8_t *chunk, uint32_t len, uint32_t *positions, uint64_t cap) {
        if (len == 0)
                return 0;
        const uint32_t entries = rd16(chunk);
        const uint8_t *end     = chunk + len - (size_t)entries * 8;
        const uint8_t *p       = chunk + 2;
        uint64_t       n       = 0;
        uint32_t       fr[32];
        while (p < end) {
                (void)vb_get(&p);
                (void)vb_get(&p);
                const uint32_t cnt = *p++;
                if (cnt == 0 || cnt > 32)
                        return -1;
                for (uint32_t i = 0; i + 1 < cnt; ++i)
                        (void)vb_get(&p);
                for (uint32_t i = 0; i < cnt; ++i)
                        fr[i] = vb_get(&p);
                for (uint32_t i = 0; i < cnt; ++i) {
                        uint32_t pos         = 0;
                        uint8_t  payloadSize = 0;
                        for (uint32_t h = 0; h < fr[i]; ++h) {
                                const uint32_t step = vb_get(&p);
                                if (step & 1u)
                                        payloadSize = *p++;
                                pos += step >> 1;
                                p += payloadSize;
                                if (n < cap)
                                        positions[n] = pos;
                                ++n;
                        }
                }
        }
        return p == end ? (int64_t)n : -1;
}

/* LUCENE: the term's hits live in hits.data at the chunk header's hitsDataOffset (lucene_codec.cpp:401-513 refill_hits, :767-856
 * materialize_hits; encoder :200-330): blocks of 128 hits = int-block(position deltas) int-block(payload sizes)
 * varbyte(payload bytes) payloads, then the last (sumHits % 128) hits as varbyte((delta << 1) | payloadSizeChanged) [u8 size] followed
 * by their payload bytes.  Position deltas restart at every document; document i owns the next freq(i) hits of the stream.
 * freqs[] = the term's per-document freqs (orc_decode_lucene).  Returns the number of positions or -1. */
int64_t orc_positions_lucene(const uint8_t *chunk, uint32_t len, const uint8_t *hits, const uint32_t *freqs, uint32_t documents, uint32_t *positions,
                             uint64_t cap) {
        if (len == 0)
                return 0;
        const uint32_t hitsOff = rd32(chunk), sumHits = rd32(chunk + 4);
        const uint8_t *p       = hits + hitsOff;
        uint32_t       delta[128], psize[128];
        uint32_t       inBlock = 0, blockFill = 0; /* cursor inside the decoded block / its number of hits */
        uint64_t       consumed = 0, n = 0;
        int            tail = 0;
        uint8_t        tailPayload = 0;
        for (uint32_t d = 0; d < documents; ++d) {
                uint32_t pos = 0;
                for (uint32_t h = 0; h < freqs[d]; ++h) {
                        if (inBlock == blockFill) { /* refill */
                                const uint64_t left = (uint64_t)sumHits - consumed;
                                if (left == 0)
                                        return -1;
                                if (left >= 128) {
                                        p = orc_ints_decode(p, delta);
                                        p = orc_ints_decode(p, psize);
                                        const uint32_t plen = vb_get(&p);
                                        p += plen;
                                        blockFill = 128;
                                        tail      = 0;
                                } else {
                                        blockFill = (uint32_t)left;
                                        tail      = 1;
                                        for (uint32_t i = 0; i < blockFill; ++i) {
                                                const uint32_t v = vb_get(&p);
                                                if (v & 1u)
                                                        tailPayload = *p++;
                                                delta[i] = v >> 1;
                                                psize[i] = tailPayload;
                                        }
                                }
                                inBlock = 0;
                        }
                        pos += delta[inBlock++];
                        ++consumed;
                        if (n < cap)
                                positions[n] = pos;
                        ++n;
                }
        }
        (void)tail;
        return (int64_t)n;
}

/* ---- BM25: IndexSourcesCollectionBM25Scorer::Scorer::idf (similarity.h:179-181, float arithmetic) and score (:228-235) ---- */
double orc_bm25_idf(uint32_t docFreq, uint64_t docsCnt) {
        const float a = (float)(docsCnt - docFreq) + 0.5f;
        const float b = (float)docFreq + 0.5f;
        return (double)logf(1.0f + a / b);
}
float orc_bm25_score(double idf, uint16_t freq) {
        const float f = (float)freq;
        return (float)(idf * (double)f / (double)(f + 1.2f));
}

/* ---- docset algebra + structural scoring ----
 * node layout == trn_qnode of include/trinity_b200.h (kind 0 TERM, 1 AND, 2 OR, 3 NOT(req, excl), 4 OPTIONAL(main, opt), 5 SOME(children, min = term), 6 PHRASE(terms in order)).
 * Matching: Conjuction / Disjunction / Filter / Optional next()/advance() semantics (docset_iterators.cpp:282-677,
 * docset_iterators.h:174-206).  Scoring: the IteratorScorer wrappers (docset_iterators_scorers.cpp:8-242): a conjunction sums all
 * children, a disjunction sums the children positioned on the document, a filter scores its required side only, an optional adds
 * its optional side when that is on the document.  Per-posting score = float, accumulated in double (docset_spans.cpp:735). */
typedef struct {
        uint8_t  kind, nchildren;
        uint16_t first_child;
        uint32_t term;
        double   weight;
} orc_node;
typedef struct {
        uint32_t documents, chunk_off, chunk_len;
} orc_term;

typedef struct {
        int             codec;
        const uint8_t * index;
        const orc_term *terms;
        const orc_node *nodes;
        uint32_t        ndocs;
        int             scored;
        const uint8_t * hits; /* Lucene hits.data (phrases only); NULL otherwise */
} orc_ctx;

static int orc_eval_node(const orc_ctx *c, uint32_t i, uint8_t *m, double *s) {
        const orc_node *X = &c->nodes[i];
        const size_t    n = (size_t)c->ndocs + 1;
        memset(m, 0, n);
        if (c->scored)
                memset(s, 0, n * sizeof(double));
        if (X->kind == 0) {
                if (X->term == 0xffffffffu)
                        return 0;
                const orc_term *t   = &c->terms[X->term];
                uint32_t *      ids = (uint32_t *)malloc(((size_t)t->documents + 1) * 4), *fr = (uint32_t *)malloc(((size_t)t->documents + 1) * 4);
                const int64_t   k   = c->codec == 0 ? orc_decode_google(c->index + t->chunk_off, t->chunk_len, ids, fr, t->documents)
                                                : orc_decode_lucene(c->index + t->chunk_off, t->chunk_len, t->documents, ids, fr, t->documents);
                if (k != (int64_t)t->documents) {
                        free(ids);
                        free(fr);
                        return -1;
                }
                for (int64_t j = 0; j < k; ++j) {
                        if (ids[j] > c->ndocs)
                                continue;
                        m[ids[j]] = 1;
                        if (c->scored)
                                s[ids[j]] = (double)orc_bm25_score(X->weight, (uint16_t)fr[j]);
                }
                free(ids);
                free(fr);
                return 0;
        }
        if (X->kind == 6) {
                /* PHRASE == Phrase::consider_phrase_match (docset_iterators.cpp:66-158) + the Phrase scorer (docset_iterators_scorers.cpp:195-228):
                 * children are the terms in order; every non-zero position q of the first term with term j at q + j for all j is one match;
                 * score = score(matchCnt, sum of the terms' idf).  Positions come from orc_positions_*. */
                const uint32_t k = X->nchildren;
                if (k < 2 || k > 16)
                        return -1;
                uint32_t *ids[16], *fr[16], *pos[16];
                uint64_t *start[16]; /* start[j][i] = index of document i's first position of term j */
                int       ok = 1;
                double    w  = 0;
                for (uint32_t j = 0; j < k; ++j)
                        ids[j] = fr[j] = pos[j] = NULL, start[j] = NULL;
                for (uint32_t j = 0; j < k && ok; ++j) {
                        const orc_node *T = &c->nodes[X->first_child + j];
                        if (T->kind != 0 || T->term == 0xffffffffu) {
                                ok = 0; /* a phrase with an unknown term matches nothing */
                                break;
                        }
                        w += T->weight;
                        const orc_term *t = &c->terms[T->term];
                        ids[j]            = (uint32_t *)malloc(((size_t)t->documents + 1) * 4);
                        fr[j]             = (uint32_t *)malloc(((size_t)t->documents + 1) * 4);
                        start[j]          = (uint64_t *)malloc(((size_t)t->documents + 1) * 8);
                        const int64_t nd  = c->codec == 0 ? orc_decode_google(c->index + t->chunk_off, t->chunk_len, ids[j], fr[j], t->documents)
                                                        : orc_decode_lucene(c->index + t->chunk_off, t->chunk_len, t->documents, ids[j], fr[j], t->documents);
                        if (nd != (int64_t)t->documents)
                                return -1;
                        uint64_t tot = 0;
                        for (uint32_t i = 0; i < t->documents; ++i) {
                                start[j][i] = tot;
                                tot += fr[j][i];
                        }
                        start[j][t->documents] = tot;
                        pos[j]                 = (uint32_t *)malloc((size_t)(tot + 1) * 4);
                        const int64_t np       = c->codec == 0 ? orc_positions_google(c->index + t->chunk_off, t->chunk_len, pos[j], tot)
                                                         : orc_positions_lucene(c->index + t->chunk_off, t->chunk_len, c->hits, fr[j], t->documents, pos[j], tot);
                        if (np != (int64_t)tot)
                                return -1;
                }
                if (ok) {
                        const orc_term *t0 = &c->terms[c->nodes[X->first_child].term];
                        uint32_t        cur[16];
                        for (uint32_t j = 0; j < k; ++j)
                                cur[j] = 0;
                        for (uint32_t i = 0; i < t0->documents; ++i) {
                                const uint32_t d = ids[0][i];
                                if (d > c->ndocs)
                                        continue;
                                uint32_t at[16];
                                int      all = 1;
                                at[0]        = i;
                                for (uint32_t j = 1; j < k && all; ++j) {
                                        const orc_term *tj = &c->terms[c->nodes[X->first_child + j].term];
                                        while (cur[j] < tj->documents && ids[j][cur[j]] < d)
                                                ++cur[j];
                                        if (cur[j] >= tj->documents || ids[j][cur[j]] != d)
                                                all = 0;
                                        at[j] = cur[j];
                                }
                                if (!all)
                                        continue;
                                uint32_t cnt = 0;
                                for (uint64_t a = start[0][i]; a < start[0][i + 1]; ++a) {
                                        const uint32_t q = pos[0][a];
                                        if (!q)
                                                continue;
                                        int hit = 1;
                                        for (uint32_t j = 1; j < k && hit; ++j) {
                                                int found = 0;
                                                for (uint64_t b = start[j][at[j]]; b < start[j][at[j] + 1] && !found; ++b)
                                                        found = pos[j][b] == q + j;
                                                hit = found;
                                        }
                                        cnt += (uint32_t)hit;
                                }
                                if (cnt) {
                                        m[d] = 1;
                                        if (c->scored)
                                                s[d] = (double)orc_bm25_score(w, (uint16_t)cnt);
                                }
                        }
                }
                for (uint32_t j = 0; j < k; ++j) {
                        free(ids[j]);
                        free(fr[j]);
                        free(pos[j]);
                        free(start[j]);
                }
                return 0;
        }
        uint8_t *cm = (uint8_t *)malloc(n);
        double * cs = c->scored ? (double *)malloc(n * sizeof(double)) : NULL;
        int      rc = 0;
        if (X->kind == 5) {
                /* SOME == DisjunctionSome (docset_iterators.cpp:679-811): >= X->term children match; the children that match score
                 * (docset_iterators_scorers.cpp:38-56) */
                uint8_t *cnt = (uint8_t *)calloc(n, 1);
                for (uint32_t k = 0; k < X->nchildren && rc == 0; ++k) {
                        rc = orc_eval_node(c, X->first_child + k, cm, cs);
                        for (size_t d = 0; d < n && rc == 0; ++d)
                                if (cm[d]) {
                                        if (cnt[d] < 255)
                                                ++cnt[d];
                                        if (cs) s[d] += cs[d];
                                }
                }
                for (size_t d = 0; d < n; ++d) {
                        m[d] = cnt[d] >= X->term;
                        if (cs && !m[d])
                                s[d] = 0;
                }
                free(cnt);
                free(cm);
                free(cs);
                return rc;
        }
        for (uint32_t k = 0; k < X->nchildren && rc == 0; ++k) {
                rc = orc_eval_node(c, X->first_child + k, cm, cs);
                if (rc)
                        break;
                for (size_t d = 0; d < n; ++d) {
                        switch (X->kind) {
                                case 1: /* AND */
                                        if (k == 0) {
                                                m[d] = cm[d];
                                                if (cs) s[d] = cs[d];
                                        } else {
                                                m[d] = m[d] && cm[d];
                                                if (cs) s[d] += cs[d];
                                        }
                                        break;
                                case 2: /* OR */
                                        if (cm[d]) {
                                                m[d] = 1;
                                                if (cs) s[d] += cs[d];
                                        }
                                        break;
                                case 3: /* NOT: req, then excl */
                                        if (k == 0) {
                                                m[d] = cm[d];
                                                if (cs) s[d] = cs[d];
                                        } else if (cm[d])
                                                m[d] = 0;
                                        break;
                                default: /* OPTIONAL: main, then opt */
                                        if (k == 0) {
                                                m[d] = cm[d];
                                                if (cs) s[d] = cs[d];
                                        } else if (cm[d] && cs)
                                                s[d] += cs[d];
                                        break;
                        }
                }
        }
        if (cs)
                for (size_t d = 0; d < n; ++d)
                        if (!m[d])
                                s[d] = 0;
        free(cm);
        free(cs);
        return rc;
}

/* DocsSetIterators::cost (docset_iterators.cpp:10-64) */
static uint64_t orc_cost(const orc_ctx *c, uint32_t i) {
        const orc_node *X = &c->nodes[i];
        if (X->kind == 0)
                return X->term == 0xffffffffu ? 0 : c->terms[X->term].documents;
        if (X->kind == 3 || X->kind == 4)
                return orc_cost(c, X->first_child);
        if (X->kind == 6) { /* Phrase: its lead term */
                uint64_t r = ~0ull;
                for (uint32_t k = 0; k < X->nchildren; ++k) {
                        const uint64_t v = orc_cost(c, X->first_child + k);
                        r                = v < r ? v : r;
                }
                return r;
        }
        if (X->kind == 5) { /* DisjunctionSome::cost_: the (size - min + 1) cheapest children */
                uint64_t v[256], r = 0;
                for (uint32_t k = 0; k < X->nchildren; ++k)
                        v[k] = orc_cost(c, X->first_child + k);
                for (uint32_t a = 0; a < X->nchildren; ++a)
                        for (uint32_t b = a + 1; b < X->nchildren; ++b)
                                if (v[b] < v[a]) {
                                        const uint64_t t = v[a];
                                        v[a]             = v[b];
                                        v[b]             = t;
                                }
                for (uint32_t k = 0; k + X->term < X->nchildren + 1u; ++k)
                        r += v[k];
                return r;
        }
        uint64_t r = X->kind == 1 ? ~0ull : 0;
        for (uint32_t k = 0; k < X->nchildren; ++k) {
                const uint64_t v = orc_cost(c, X->first_child + k);
                r                = X->kind == 1 ? (v < r ? v : r) : r + v;
        }
        return r;
}

/* exec_query's general path (exec.cpp:1083-1345) for DocumentsOnly (scored = 0) / AccumulatedScoreScheme (scored = 1).
 * match[d] = 1 for every matched docID d in 1..ndocs, score[d] = accumulated score.  Includes build_span's behaviour for a root
 * Filter over a disjunction whose excluded side is not costlier (exec.cpp:488-501 + docset_spans.cpp:98-111,681-694: the
 * disjunction spans ignore `min`, so the exclusion is not applied) — observed on the reference, see tests. */
int orc_exec2(int codec, const uint8_t *index, const uint8_t *hits, const orc_term *terms, const orc_node *nodes, uint32_t ndocs, int scored, uint8_t *match,
              double *score);
int orc_exec(int codec, const uint8_t *index, const orc_term *terms, const orc_node *nodes, uint32_t ndocs, int scored, uint8_t *match, double *score) {
        return orc_exec2(codec, index, NULL, terms, nodes, ndocs, scored, match, score);
}
/* hits: Lucene hits.data (needed by phrase nodes on the Lucene codec; Google keeps its hits inline) */
int orc_exec2(int codec, const uint8_t *index, const uint8_t *hits, const orc_term *terms, const orc_node *nodes, uint32_t ndocs, int scored, uint8_t *match,
              double *score) {
        orc_ctx  c    = {codec, index, terms, nodes, ndocs, scored, hits};
        uint32_t root = 0, cur = 0;
        int      traversed = 0;
        while (nodes[cur].kind == 3 && orc_cost(&c, nodes[cur].first_child + 1u) <= orc_cost(&c, nodes[cur].first_child)) {
                cur       = nodes[cur].first_child;
                traversed = 1;
        }
        if (traversed && nodes[cur].kind == 2)
                root = cur;
        return orc_eval_node(&c, root, match, score);
}
