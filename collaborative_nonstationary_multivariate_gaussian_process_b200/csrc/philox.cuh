// Counter-based standard-normal noise (Philox4x32-10 + Box-Muller in float32, then widened to double: the reference
// draws float32 normals and casts them, code/utils.py:123,226,234, quirk q2).  A draw is a pure function of
// (seed, stream, sample, global row id, column), so runs sharded over 1/2/4/8 ranks see the same noise.
#pragma once
#include <stdint.h>

struct NoiseKey {
    unsigned long long seed;
    unsigned long long stream;   // (step << 8) | kind
    // optional DEVICE step counter: the effective stream id is stream | (*step_dev << 8).  A step captured in a CUDA graph
    // passes stream = kind and bumps the counter inside the graph, so every replay draws fresh noise -- the same noise
    // an eager run with stream = (step << 8) | kind draws.
    const unsigned long long* step_dev;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// four N(0,1) float32 values for (sample, row gid, column quad q4)
__device__ __forceinline__ void philox_normal4(NoiseKey key, unsigned int sample, unsigned long long gid, unsigned int q4,
                                               float z[4]) {
    uint32_t u[4];
    const unsigned long long stream = key.step_dev ? (key.stream | (__ldg(key.step_dev) << 8)) : key.stream;
    // counter = (row gid lo, row gid hi, sample, column quad); key = seed words xor-ed with the stream id
    philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), sample, q4, (uint32_t)key.seed ^ (uint32_t)stream,
                  (uint32_t)(key.seed >> 32) ^ (uint32_t)(stream >> 32), u);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)(u[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0,1), 24 bits
        const float u2 = ((float)(u[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(6.28318530717958647692f * u2, &sn, &cs);
        z[2 * h] = r * cs;
        z[2 * h + 1] = r * sn;
    }
}
