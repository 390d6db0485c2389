// Saru PRNG (Afshar, Schmid, Pishevar, Worley, Comput. Phys. Commun. 184 (2013) 1119) as used by the reference
// (third_party/saruprng.cuh:257-280 three-seed constructor, :196-213,:339-351 stepping/output, :361-365 float
// conversion, :115-128 Box-Muller). Restated; the constants are the algorithm.
#pragma once
#include <cfloat>
#include <cstdint>

namespace ub200 {

struct Saru {
  uint32_t lcg, weyl;
  __device__ __forceinline__ Saru(uint32_t s1, uint32_t s2, uint32_t s3) {
    s3 ^= (s1 << 7) ^ (s2 >> 6);
    s2 += (s1 >> 4) ^ (s3 >> 15);
    s1 ^= (s2 << 9) + (s3 << 8);
    s3 ^= 0xA5366B4Du * ((s2 >> 11) ^ (s1 << 1));
    s2 += 0x72BE1579u * ((s1 << 4) ^ (s3 >> 16));
    s1 ^= 0x3F38A6EDu * ((s3 >> 5) ^ (uint32_t)(((int32_t)s2) >> 22));
    s2 += s1 * s3;
    s1 += s3 ^ (s2 >> 2);
    s2 ^= (uint32_t)(((int32_t)s2) >> 17);
    lcg = 0x79dedea3u * (s1 ^ (uint32_t)(((int32_t)s1) >> 14));
    weyl = (lcg + s2) ^ (uint32_t)(((int32_t)lcg) >> 8);
    lcg = lcg + (weyl * (weyl ^ 0xdddf97f5u));
    weyl = 0xABCB96F7u + (weyl >> 1);
  }
  __device__ __forceinline__ uint32_t u32() {
    lcg = 0x4beb5d59u * lcg + 0x2600e1f7u;
    weyl = weyl + 0x8009d14bu + ((uint32_t)(((int32_t)weyl) >> 31) & 0xda879addu);
    const uint32_t v = (lcg ^ (lcg >> 26)) + weyl;
    return (v ^ (v >> 20)) * 0x6957f5a7u;
  }
  __device__ __forceinline__ float f() { return ((int32_t)(u32() >> 1)) * (1.0f / 2147483648.0f); }
  // first component of the Box-Muller pair gf(mean=0, std)
  __device__ __forceinline__ float gaussX(float std) {
    float u0;
    do { u0 = f(); } while (u0 <= FLT_MIN);
    const float u1 = f();
    const float r = sqrtf(-2.0f * logf(u0));
    const float theta = 6.283185307179586f * u1;
    return __fmul_rn(r * sinf(theta), std);
  }
  // full Box-Muller pair gf(0, std): float arithmetic even in double precision builds (saruprng.cuh:115-128)
  __device__ __forceinline__ float2 gauss2(float std) {
    float u0;
    do { u0 = f(); } while (u0 <= FLT_MIN);
    const float u1 = f();
    const float r = sqrtf(-2.0f * logf(u0));
    const float theta = 6.283185307179586f * u1;
    // gf() is "(r sin, r cos) * std + mean" with mean = 0 (saruprng.cuh:127): the product is rounded on its own (an
    // fma with a zero addend), it must never be contracted into the caller's accumulation
    return make_float2(__fmul_rn(r * sinf(theta), std), __fmul_rn(r * cosf(theta), std));
  }
};

} // namespace ub200
