/* TEST INFRASTRUCTURE - C restatement of the Saru PRNG (Y. Afshar, F. Schmid, A. Pishevar, S. Worley,
 * "Exploiting seeding of random number generators for efficient domain decomposition parallelization of
 * dissipative particle dynamics", Comput. Phys. Commun. 184 (2013) 1119) as vendored by the reference at
 * third_party/saruprng.cuh: three-seed constructor :257-280, single-step LCG + offset-Weyl advance
 * :196-213,229-233, output mix :339-351, float conversion :361-365, Box-Muller gf :115-128.
 * The constants are the algorithm; the code structure is ours.
 */
#ifndef UAMMD_B200_ORACLE_SARU_H
#define UAMMD_B200_ORACLE_SARU_H
#include <float.h>
#include <math.h>
#include <stdint.h>

typedef struct {
  uint32_t lcg;
  uint32_t weyl;
} orc_saru;

static inline int32_t orc_sar(uint32_t v, int s) { return ((int32_t)v) >> s; } /* arithmetic shift */

/* final state churn shared by all constructors (saruprng.cuh:275-278) */
static inline orc_saru orc_saru_finish(uint32_t a, uint32_t b) {
  orc_saru r;
  r.lcg = 0x79dedea3u * (a ^ (uint32_t)orc_sar(a, 14));
  r.weyl = (r.lcg + b) ^ (uint32_t)orc_sar(r.lcg, 8);
  r.lcg = r.lcg + (r.weyl * (r.weyl ^ 0xdddf97f5u));
  r.weyl = 0xABCB96F7u + (r.weyl >> 1);
  return r;
}

static inline orc_saru orc_saru_seed3(uint32_t s1, uint32_t s2, uint32_t s3) {
  s3 ^= (s1 << 7) ^ (s2 >> 6);
  s2 += (s1 >> 4) ^ (s3 >> 15);
  s1 ^= (s2 << 9) + (s3 << 8);
  s3 ^= 0xA5366B4Du * ((s2 >> 11) ^ (s1 << 1));
  s2 += 0x72BE1579u * ((s1 << 4) ^ (s3 >> 16));
  s1 ^= 0x3F38A6EDu * ((s3 >> 5) ^ (uint32_t)orc_sar(s2, 22));
  s2 += s1 * s3;
  s1 += s3 ^ (s2 >> 2);
  s2 ^= (uint32_t)orc_sar(s2, 17);
  return orc_saru_finish(s1, s2);
}

/* two-seed constructor (saruprng.cuh:236-251) */
static inline orc_saru orc_saru_seed2(uint32_t seed1, uint32_t seed2) {
  seed2 += seed1 << 16;
  seed1 += seed2 << 11;
  seed2 += (uint32_t)orc_sar(seed1, 7);
  seed1 ^= (uint32_t)orc_sar(seed2, 3);
  seed2 *= 0xA5366B4Du;
  seed2 ^= seed2 >> 10;
  seed2 ^= (uint32_t)orc_sar(seed2, 19);
  seed1 += seed2 ^ 0x6d2d4e11u;
  return orc_saru_finish(seed1, seed2);
}

static inline uint32_t orc_saru_u32(orc_saru *r) {
  r->lcg = 0x4beb5d59u * r->lcg + 0x2600e1f7u;                                /* LCG, one step */
  r->weyl = r->weyl + 0x8009d14bu + ((uint32_t)orc_sar(r->weyl, 31) & 0xda879addu); /* offset Weyl, one step */
  uint32_t v = (r->lcg ^ (r->lcg >> 26)) + r->weyl;
  return (v ^ (v >> 20)) * 0x6957f5a7u;
}

static inline float orc_saru_f(orc_saru *r) { return ((int32_t)(orc_saru_u32(r) >> 1)) * (1.0f / 2147483648.0f); }

/* Box-Muller pair, float arithmetic even in double builds (saruprng.cuh:115-128) */
static inline void orc_saru_gf(orc_saru *r, float mean, float std, float out[2]) {
  const float pi2 = (float)(2.0 * M_PI);
  float u0;
  do {
    u0 = orc_saru_f(r);
  } while (u0 <= FLT_MIN);
  const float u1 = orc_saru_f(r);
  const float rad = sqrtf(-2.0f * logf(u0));
  const float theta = pi2 * u1;
  out[0] = (rad * sinf(theta)) * std + mean;
  out[1] = (rad * cosf(theta)) * std + mean;
}
/* Box-Muller pair gd: float uniforms, float log/sqrt/sin/cos, double products (saruprng.cuh:130-143) */
static inline void orc_saru_gd(orc_saru *r, double mean, double std, double out[2]) {
  const double pi2 = 2.0 * M_PI;
  double u0;
  do {
    u0 = orc_saru_f(r);
  } while (u0 <= DBL_MIN);
  const double u1 = orc_saru_f(r);
  const double rad = sqrtf((float)(-2.0 * logf((float)u0)));
  const double theta = pi2 * u1;
  out[0] = (rad * sinf((float)theta)) * std + mean;
  out[1] = (rad * cosf((float)theta)) * std + mean;
}
#endif
