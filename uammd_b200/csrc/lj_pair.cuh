// LJ pair body and warp reductions shared by the cell traversal (pair_lj.cu) and the column traversal (lj_column.cu).
#pragma once
#include "pair_common.cuh"

namespace ub200 {

struct Acc {
  float fx, fy, fz, e, v;
};

// Packed single precision (sm_100: add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2, two IEEE fp32 operations per issued
// instruction; a {s, s} pair built from one register is folded by ptxas into the instruction's scalar-broadcast operand).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// One LJ pair, branch free. r2 is a non negative float, so its bit pattern orders like an unsigned integer:
// (bits(r2) - 1) < (bits(rc2) - 1) <=> 0 < r2 < rc2 (r2 == 0 wraps to 0xffffffff) - two integer-pipe
// instructions instead of two FSETPs. Out-of-range pairs get r2 = +inf, hence 1/r2 = 0 and a zero force.
// The reciprocal is MUFU.RCP (1 ulp); with u = sigma2/r2: |F|/r = epsDivSigma2 (24 - 48 u^3) u^4.
template <bool ENERGY, bool VIRIAL>
__device__ __forceinline__ void ljPair(float dx, float dy, float dz, const LJPar &p, uint32_t rc2bitsm1, Acc &a) {
  const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
  const bool in = (__float_as_uint(r2) - 1u) < rc2bitsm1;
  const float r2s = in ? r2 : __int_as_float(0x7f800000);
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2s));
  const float u = p.sigma2 * inv;
  const float u2 = u * u;
  const float u3 = u2 * u;
  const float fm = (p.epsDivSigma2 * __fmaf_rn(-48.0f, u3, 24.0f)) * (u2 * u2);
  a.fx = __fmaf_rn(fm, dx, a.fx);
  a.fy = __fmaf_rn(fm, dy, a.fy);
  a.fz = __fmaf_rn(fm, dz, a.fz);
  if (ENERGY) a.e += in ? 0.5f * (p.epsDivSigma2 * p.sigma2 * 4.0f * u3 * (u3 - 1.0f) - p.shift) : 0.0f;
  if (VIRIAL) a.v += in ? fm * r2 : 0.0f;
}

// Force-only, single-type pair with the parameters folded into two constants: |F|/r = inv^4 (c24 - c48 inv^3), inv = 1/r2,
// c24 = 24 eps sigma^6, c48 = 48 eps sigma^12 (the same function as ljPair: epsDivSigma2 (24 - 48 u^3) u^4 with u = sigma2 inv;
// two multiplications less per pair, roundings differ in the last bits only).
struct LJFold {
  float c24, c48;
  uint32_t rcb;
};
__device__ __forceinline__ LJFold foldLJ(const LJPar &p) {
  const float s4 = p.sigma2 * p.sigma2, s8 = s4 * s4;
  LJFold f;
  f.c24 = 24.0f * p.epsDivSigma2 * s8;
  f.c48 = 48.0f * p.epsDivSigma2 * s8 * s4 * p.sigma2;
  f.rcb = __float_as_uint(p.cutOff2) - 1u;
  return f;
}
__device__ __forceinline__ void ljPairFolded(float dx, float dy, float dz, const LJFold &p, Acc &a) {
  const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
  const bool in = (__float_as_uint(r2) - 1u) < p.rcb; // 0 < r2 < rc2 (see ljPair)
  const float r2s = in ? r2 : __int_as_float(0x7f800000);
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2s));
  const float i2 = inv * inv;
  const float fm = (i2 * i2) * __fmaf_rn(-p.c48, i2 * inv, p.c24);
  a.fx = __fmaf_rn(fm, dx, a.fx);
  a.fy = __fmaf_rn(fm, dy, a.fy);
  a.fz = __fmaf_rn(fm, dz, a.fz);
}

// Sum the 2x3 force components (and optionally energy/virial) of two home particles over the warp.
// First exchange across lane halves so that each half carries one particle, then a 4 level butterfly.
__device__ __forceinline__ void reducePair(Acc &a0, Acc &a1, int lane, bool ev) {
  const bool hi = lane & 16;
  // lanes 0-15 keep particle 0, lanes 16-31 keep particle 1
  float sx = hi ? a0.fx : a1.fx, sy = hi ? a0.fy : a1.fy, sz = hi ? a0.fz : a1.fz;
  float kx = hi ? a1.fx : a0.fx, ky = hi ? a1.fy : a0.fy, kz = hi ? a1.fz : a0.fz;
  kx += __shfl_xor_sync(0xffffffffu, sx, 16);
  ky += __shfl_xor_sync(0xffffffffu, sy, 16);
  kz += __shfl_xor_sync(0xffffffffu, sz, 16);
  float se = 0.f, sv = 0.f, ke = 0.f, kv = 0.f;
  if (ev) {
    se = hi ? a0.e : a1.e; sv = hi ? a0.v : a1.v;
    ke = hi ? a1.e : a0.e; kv = hi ? a1.v : a0.v;
    ke += __shfl_xor_sync(0xffffffffu, se, 16);
    kv += __shfl_xor_sync(0xffffffffu, sv, 16);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    kx += __shfl_xor_sync(0xffffffffu, kx, o);
    ky += __shfl_xor_sync(0xffffffffu, ky, o);
    kz += __shfl_xor_sync(0xffffffffu, kz, o);
    if (ev) {
      ke += __shfl_xor_sync(0xffffffffu, ke, o);
      kv += __shfl_xor_sync(0xffffffffu, kv, o);
    }
  }
  // result for particle 0 in lane 0, particle 1 in lane 16 (stored in a0)
  a0.fx = kx; a0.fy = ky; a0.fz = kz; a0.e = ke; a0.v = kv;
}

} // namespace ub200
