// Spectral operator of the PSE far field (Hasimoto-split RPY Green's function), applied inside the fused z pass.
#pragma once
#include "fft.cuh"
#include "saru.cuh"

namespace ub200 {

// ------------------------------------------------------------------------------------------------------------
// Far field spectral operator
// ------------------------------------------------------------------------------------------------------------
template <class T> struct PseSpectralOp {
  using C = typename Vec2<T>::type;
  int nx, ny, nz, nkx;
  T kfx, kfy, kfz; // 2 pi / L (waveNumberToWaveVector, PSE/utils.cuh:41-44)
  T shear, rh, vis, split, eta, nTot;
  int deterministic, noise;
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
  __device__ __forceinline__ bool nyquist(int ix, int iy, int iz) const { // FarField.cuh:183-219
    const bool nxq = (ix == nx - ix) && (nx % 2 == 0);
    const bool nyq = (iy == ny - iy) && (ny % 2 == 0);
    const bool nzq = (iz == nz - iz) && (nz % 2 == 0);
    return (nxq && iy == 0 && iz == 0) || (nxq && nyq && iz == 0) || (ix == 0 && nyq && iz == 0) ||
           (nxq && iy == 0 && nzq) || (ix == 0 && iy == 0 && nzq) || (ix == 0 && nyq && nzq) || (nxq && nyq && nzq);
  }
  __device__ __forceinline__ void drawNoise(uint32_t id, C &a, C &b, C &c) const { // generateNoise :161-177
    Saru rng(id, seed1, seed2Dev ? *seed2Dev : seed2);
    const float sc = (float)(T(0.707106781186547) * noisePrefactor);
    float2 g = rng.gauss2(sc); a = mk2<T>((T)g.x, (T)g.y);
    g = rng.gauss2(sc); b = mk2<T>((T)g.x, (T)g.y);
    g = rng.gauss2(sc); c = mk2<T>((T)g.x, (T)g.y);
  }
  // greensFunction (FarField.cuh:85-119); (kx, ky, kz) is the unsheared (NUFFT) wave vector, kyE the sheared ky
  __device__ __forceinline__ T greens(T kx, T ky, T kz, T kyE) const {
    const T kN2 = kx * kx + ky * ky + kz * kz;
    if (kN2 == T(0)) return T(0);
    const T kE2 = kx * kx + kyE * kyE + kz * kz;
    const T kmod = sqrt(kE2);
    const T invk2 = T(1.0) / kE2;
    const T sink = sin(kmod * rh);
    const T kEw = kE2 / (T(4.0) * split * split);
    const T kNu = kN2 / (T(4.0) * split * split);
    const T tau = eta * kNu - kEw;
    const T hashimoto = (T(1.0) + kEw) * exp(tau) / kE2;
    T B = sink * sink * invk2 * hashimoto / (vis * rh * rh);
    B /= nTot;
    return B;
  }

  __device__ __forceinline__ void operator()(int ix, int iy, int iz, C &vx, C &vy, C &vz) const {
    iy += yOff;
    if (ix == 0 && iy == 0 && iz == 0) { vx = vy = vz = mk2<T>(T(0), T(0)); return; }
    const T kx = kfx * (T)fold(ix, nx), ky = kfy * (T)fold(iy, ny), kz = kfz * (T)fold(iz, nz);
    const T kyE = ky - shear * kx; // shearWaveVector (PSE/utils.cuh:36-39)
    const T B = greens(kx, ky, kz, kyE);
    const T invk2 = T(1.0) / (kx * kx + kyE * kyE + kz * kz);
    auto project = [&](T f0, T f1, T f2, T &o0, T &o1, T &o2) { // projectFourier (FarField.cuh:53-73)
      const T kf = (kx * f0 + kyE * f1 + kz * f2) * invk2;
      o0 = f0 - kx * kf; o1 = f1 - kyE * kf; o2 = f2 - kz * kf;
    };
    C ox = mk2<T>(T(0), T(0)), oy = ox, oz = ox;
    if (deterministic) {
      T a0, a1, a2, b0, b1, b2;
      project(B * vx.x, B * vy.x, B * vz.x, a0, a1, a2);
      project(B * vx.y, B * vy.y, B * vz.y, b0, b1, b2);
      ox = mk2<T>(a0, b0); oy = mk2<T>(a1, b1); oz = mk2<T>(a2, b2);
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
        project(n0.x, n1.x, n2.x, a0, a1, a2);
        project(n0.y, n1.y, n2.y, b0, b1, b2);
        ox.x += Bsq * a0; ox.y += Bsq * b0; oy.x += Bsq * a1; oy.y += Bsq * b1; oz.x += Bsq * a2; oz.y += Bsq * b2;
      }
      // what the conjugate partner adds here (stored twice only on the kx = 0 and kx = nx/2 planes); the
      // reference does this with a second non-atomic "+=" from another thread, this is the race-free sum
      if (ix == 0 || ix == nx - ix) {
        const int cy = (iy > 0) * (ny - iy), cz = (iz > 0) * (nz - iz);
        if (!(cy == iy && cz == iz) && generates(ix, cy, cz) && !nyquist(ix, cy, cz)) {
          C n0, n1, n2;
          drawNoise((uint32_t)(ix + nkx * (cy + ny * cz)), n0, n1, n2);
          T a0, a1, a2, b0, b1, b2;
          project(n0.x, n1.x, n2.x, a0, a1, a2);
          project(-n0.y, -n1.y, -n2.y, b0, b1, b2);
          ox.x += Bsq * a0; ox.y += Bsq * b0; oy.x += Bsq * a1; oy.y += Bsq * b1; oz.x += Bsq * a2; oz.y += Bsq * b2;
        }
      }
    }
    vx = ox; vy = oy; vz = oz;
  }
};

} // namespace ub200
