// Prefix-varbyte wire format shared by both codecs (host + device).
// Format reference: Switch/switch_compiler_aux.h:23-81 (varbyte_put32 / varbyte_get32):
//   0xxxxxxx                      7 bits
//   10xxxxxx b1                   14 bits, big-endian
//   110xxxxx lo hi                21 bits: top 5 bits in byte 0, low 16 bits little-endian
//   1110xxxx b1 b2 b3             28 bits, big-endian
//   11110000 u32le                32 bits
#pragma once
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define TRN_HD __host__ __device__ __forceinline__
#else
#define TRN_HD inline
#endif

namespace trn {

// length in bytes of the code starting with first byte b0
TRN_HD uint32_t varbyte_len(uint32_t b0) {
        if (b0 < 0x80u) return 1;
        if (b0 < 0xc0u) return 2;
        if (b0 < 0xe0u) return 3;
        if (b0 < 0xf0u) return 4;
        return 5;
}

// decode one code at p, advance p
TRN_HD uint32_t varbyte_get(const uint8_t *&p) {
        const uint32_t b0 = *p++;
        if (b0 < 0x80u) return b0;
        if (b0 < 0xc0u) {
                const uint32_t v = ((b0 & 0x3fu) << 8) | p[0];
                p += 1;
                return v;
        }
        if (b0 < 0xe0u) {
                const uint32_t v = ((b0 & 0x1fu) << 16) | p[0] | (uint32_t(p[1]) << 8);
                p += 2;
                return v;
        }
        if (b0 < 0xf0u) {
                const uint32_t v = ((b0 & 0x0fu) << 24) | (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2];
                p += 3;
                return v;
        }
        const uint32_t v = p[0] | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
        p += 4;
        return v;
}

#if !defined(__CUDA_ARCH__)
inline void varbyte_put(std::vector<uint8_t> &out, uint32_t x) {
        if (x < (1u << 7)) {
                out.push_back(uint8_t(x));
        } else if (x < (1u << 14)) {
                out.push_back(uint8_t(0x80u | (x >> 8)));
                out.push_back(uint8_t(x));
        } else if (x < (1u << 21)) {
                out.push_back(uint8_t(0xc0u | (x >> 16)));
                out.push_back(uint8_t(x));
                out.push_back(uint8_t(x >> 8));
        } else if (x < (1u << 28)) {
                out.push_back(uint8_t(0xe0u | (x >> 24)));
                out.push_back(uint8_t(x >> 16));
                out.push_back(uint8_t(x >> 8));
                out.push_back(uint8_t(x));
        } else {
                out.push_back(0xf0u);
                out.push_back(uint8_t(x));
                out.push_back(uint8_t(x >> 8));
                out.push_back(uint8_t(x >> 16));
                out.push_back(uint8_t(x >> 24));
        }
}

inline void put_u16(std::vector<uint8_t> &out, uint16_t v) {
        out.push_back(uint8_t(v));
        out.push_back(uint8_t(v >> 8));
}
inline void put_u32(std::vector<uint8_t> &out, uint32_t v) {
        out.push_back(uint8_t(v));
        out.push_back(uint8_t(v >> 8));
        out.push_back(uint8_t(v >> 16));
        out.push_back(uint8_t(v >> 24));
}
inline uint32_t get_u32(const uint8_t *p) {
        return p[0] | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
inline uint16_t get_u16(const uint8_t *p) {
        return uint16_t(p[0] | (uint16_t(p[1]) << 8));
}
#endif

} // namespace trn
