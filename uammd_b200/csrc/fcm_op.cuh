// Spectral operator of the Force Coupling Method, applied inside the fused z pass of the 3-D FFT.
#pragma once
#include "fft.cuh"
#include "saru.cuh"

namespace ub200 {

// ---- spectral operator of FCM: forceFourier2Vel (FCM_impl.cuh:375-397) + fourierBrownianNoise (:437-512) ----
template <class T> struct FcmSpectralOp {
  using C = typename Vec2<T>::type;
  int nx, ny, nz, nkx;
  T kfx, kfy, kfz;  // 2 pi / L
  T vis;
  T invNorm;        // 1 / (nx ny nz)
  int deterministic; // apply B * projector to the incoming spectrum (forces were spread)
  int noise;
  T noisePrefactor;
  uint32_t seed1, seed2;
  const uint32_t *seed2Dev = nullptr; // when set, the second seed is read from the device (captured CUDA graphs of the step)
  int yOff = 0; // slab-decomposed transform: global ky of the first local row

  __device__ __forceinline__ static int fold(int i, int n) { return i - n * (i >= (n / 2 + 1)); }

  __device__ __forceinline__ bool generates(int ix, int iy, int iz) const {
    if (ix == 0 && iy == 0 && iz == 0) return false;
    if (ix == 0 && iy == 0 && 2 * iz >= nz + 1) return false;
    if (ix == 0 && 2 * iy >= ny + 1) return false;
    return true;
  }
  // fcm_detail::isNyquistWaveNumber (FCM/utils.cuh:132-168)
  __device__ __forceinline__ bool nyquist(int ix, int iy, int iz) const {
    const bool nxq = (ix == nx - ix) && (nx % 2 == 0);
    const bool nyq = (iy == ny - iy) && (ny % 2 == 0);
    const bool nzq = (iz == nz - iz) && (nz % 2 == 0);
    return (nxq && iy == 0 && iz == 0) || (nxq && nyq && iz == 0) || (ix == 0 && nyq && iz == 0) ||
           (nxq && iy == 0 && nzq) || (ix == 0 && iy == 0 && nzq) || (ix == 0 && nyq && nzq) || (nxq && nyq && nzq);
  }
  // fcm_detail::generateNoise (FCM/utils.cuh:115-130): three float Box-Muller pairs from Saru(id, seed1, seed2)
  __device__ __forceinline__ void drawNoise(uint32_t id, C &a, C &b, C &c) const {
    Saru rng(id, seed1, seed2Dev ? *seed2Dev : seed2);
    const float sc = (float)(T(0.707106781186547) * noisePrefactor);
    float2 g = rng.gauss2(sc); a = mk2<T>((T)g.x, (T)g.y);
    g = rng.gauss2(sc); b = mk2<T>((T)g.x, (T)g.y);
    g = rng.gauss2(sc); c = mk2<T>((T)g.x, (T)g.y);
  }

  __device__ __forceinline__ void operator()(int ix, int iy, int iz, C &vx, C &vy, C &vz) const {
    iy += yOff;
    if (ix == 0 && iy == 0 && iz == 0) { vx = vy = vz = mk2<T>(T(0), T(0)); return; }
    const int fx = fold(ix, nx), fy = fold(iy, ny), fz = fold(iz, nz);
    const T kx = kfx * fx, ky = kfy * fy, kz = kfz * fz;
    const T k2 = kx * kx + ky * ky + kz * kz;
    // getGradientFourier (FCM/utils.cuh:41-51): unpaired (Nyquist) components of the gradient are zeroed
    const T dx = (fx == nx - fx) ? T(0) : kx, dy = (fy == ny - fy) ? T(0) : ky, dz = (fz == nz - fz) ? T(0) : kz;
    const T invk2 = T(1.0) / k2;
    const T B = T(1.0) / (vis * k2);
    auto project = [&](T f0, T f1, T f2, T &o0, T &o1, T &o2) { // projectFourier (FCM/utils.cuh:70-76)
      const T s = f0 * (dx * invk2) + f1 * (dy * invk2) + f2 * (dz * invk2);
      o0 = f0 - dx * s; o1 = f1 - dy * s; o2 = f2 - dz * s;
    };
    C ox = mk2<T>(T(0), T(0)), oy = ox, oz = ox;
    if (deterministic) {
      const T sc = B * invNorm;
      T a0, a1, a2, b0, b1, b2;
      project(vx.x, vy.x, vz.x, a0, a1, a2);
      project(vx.y, vy.y, vz.y, b0, b1, b2);
      ox = mk2<T>(a0 * sc, b0 * sc); oy = mk2<T>(a1 * sc, b1 * sc); oz = mk2<T>(a2 * sc, b2 * sc);
    }
    if (noise) {
      const T Bsq = sqrt(B);
      if (generates(ix, iy, iz)) {
        C n0, n1, n2;
        drawNoise((uint32_t)(ix + nkx * (iy + ny * iz)), n0, n1, n2);
        if (nyquist(ix, iy, iz)) {
          const T q = T(1.41421356237310);
          n0.x *= q; n0.y = T(0); n1.x *= q; n1.y = T(0); n2.x *= q; n2.y = T(0);
        }
        T a0, a1, a2, b0, b1, b2;
        project(n0.x * Bsq, n1.x * Bsq, n2.x * Bsq, a0, a1, a2);
        project(n0.y * Bsq, n1.y * Bsq, n2.y * Bsq, b0, b1, b2);
        ox.x += a0; ox.y += b0; oy.x += a1; oy.y += b1; oz.x += a2; oz.y += b2;
      }
      // contribution written by the conjugate partner (stored twice only on the kx = 0 and kx = nx/2 planes)
      if (ix == 0 || ix == nx - ix) {
        const int cy = (iy > 0) * (ny - iy), cz = (iz > 0) * (nz - iz);
        if (!(cy == iy && cz == iz) && generates(ix, cy, cz) && !nyquist(ix, cy, cz)) {
          C n0, n1, n2;
          drawNoise((uint32_t)(ix + nkx * (cy + ny * cz)), n0, n1, n2);
          T a0, a1, a2, b0, b1, b2;
          project(n0.x * Bsq, n1.x * Bsq, n2.x * Bsq, a0, a1, a2);
          project(-(n0.y * Bsq), -(n1.y * Bsq), -(n2.y * Bsq), b0, b1, b2);
          ox.x += a0; ox.y += b0; oy.x += a1; oy.y += b1; oz.x += a2; oz.y += b2;
        }
      }
    }
    vx = ox; vy = oy; vz = oz;
  }
};

} // namespace ub200
