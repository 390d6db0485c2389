// Hand-written batched 3-D real FFT for sm_100a (no cuFFT): shared-memory Stockham autosort passes.
//
// Replaces the cuFFT plans of FCM / PSE (cufftMakePlanMany rank 3, batch 3, interleaved real3/complex3,
// Integrator/BDHI/FCM/FCM_impl.cuh:179-234, PSE/FarField.cuh:555-603): unnormalised, forward sign -, same
// (nx/2+1) innermost Hermitian-half layout.
//
// Layout (in place): the real grid is real3 AoS [nz][ny][nxPad][3] with nxPad = 2(nx/2+1); the Fourier grid
// is complex3 AoS [nz][ny][nkx][3], nkx = nx/2+1 - byte for byte the same buffer, line by line.
//
// Passes
//   X  : R2C / C2R along the contiguous axis. Two real lines are packed into one complex transform
//        (z = a + i b; A_k = (Z_k + conj Z_{n-k})/2, B_k = -i (Z_k - conj Z_{n-k})/2), so any nx works.
//   Y,Z: complex transforms along strided axes on tiles of TX consecutive kx (TX*3 complex numbers =
//        TX*48 contiguous bytes in fp64 per line element -> full 32-byte sectors).
//   The Z pass can be fused: forward z transform -> spectral operator (Stokes / Ewald kernel, noise) ->
//   inverse z transform in one trip through shared memory.
// Every pass reads and writes each grid byte exactly once.
#pragma once
#include "common.cuh"

namespace ub200 {

template <class T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <class T> __host__ __device__ __forceinline__ typename Vec2<T>::type mk2(T x, T y) {
  typename Vec2<T>::type r;
  r.x = x; r.y = y;
  return r;
}
template <class C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <class C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <class C> __device__ __forceinline__ C cmul(C a, C b) {
  C r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR, class C> __device__ __forceinline__ C mulI(C a) {
  C r;
  if (DIR < 0) { r.x = a.y; r.y = -a.x; } else { r.x = -a.y; r.y = a.x; }
  return r;
}

constexpr int kMaxStages = 12;
struct FftAxis {
  int n;
  int nstages;
  int radix[kMaxStages];
};

// n = prod radix; prefer 4, then 2, 3, 5, 7. Returns false when n has other prime factors.
inline bool factorize(int n, FftAxis &ax) {
  ax.n = n;
  ax.nstages = 0;
  int m = n;
  while (m % 4 == 0 && ax.nstages < kMaxStages) { ax.radix[ax.nstages++] = 4; m /= 4; }
  const int primes[4] = {2, 3, 5, 7};
  for (int p : primes)
    while (m % p == 0 && ax.nstages < kMaxStages) { ax.radix[ax.nstages++] = p; m /= p; }
  return m == 1;
}

// One Stockham stage over `nf` independent transforms of length n stored at a[f*fstride + i].
// tw[j] = exp(-2 pi i j / n). DIR = -1 forward, +1 inverse (twiddles conjugated).
template <class T, int DIR>
__device__ __forceinline__ void stockhamStage(const typename Vec2<T>::type *__restrict__ a,
                                              typename Vec2<T>::type *__restrict__ b, int n, int fstride, int nf,
                                              int R, int Ns, const typename Vec2<T>::type *__restrict__ tw) {
  using C = typename Vec2<T>::type;
  const int nb = n / R;            // butterflies per transform
  const int twStep = n / (Ns * R); // twiddle index stride
  // exact small-integer division through a float reciprocal: (w + 0.5)/nb stays >= 0.5/nb away from integers
  const float invNb = 1.0f / (float)nb, invNs = 1.0f / (float)Ns;
  for (int w = threadIdx.x; w < nf * nb; w += blockDim.x) {
    const int f = __float2int_rz(((float)w + 0.5f) * invNb), j = w - f * nb;
    const int k = j - Ns * __float2int_rz(((float)j + 0.5f) * invNs);
    const C *src = a + f * fstride;
    C *dst = b + f * fstride + (j - k) * R + k;
    auto twid = [&](int idx) {
      C t = tw[idx];
      if (DIR > 0) t.y = -t.y;
      return t;
    };
    if (R == 4) {
      C v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb], v3 = src[j + 3 * nb];
      if (Ns > 1) {
        const int t1 = k * twStep;
        v1 = cmul(v1, twid(t1));
        v2 = cmul(v2, twid(2 * t1));
        v3 = cmul(v3, twid(3 * t1));
      }
      const C s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = mulI<DIR>(csub(v1, v3));
      dst[0] = cadd(s02, s13);
      dst[Ns] = cadd(d02, d13);
      dst[2 * Ns] = csub(s02, s13);
      dst[3 * Ns] = csub(d02, d13);
    } else if (R == 2) {
      C v0 = src[j], v1 = src[j + nb];
      if (Ns > 1) v1 = cmul(v1, twid(k * twStep));
      dst[0] = cadd(v0, v1);
      dst[Ns] = csub(v0, v1);
    } else if (R == 3) {
      C v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb];
      if (Ns > 1) {
        const int t1 = k * twStep;
        v1 = cmul(v1, twid(t1));
        v2 = cmul(v2, twid(2 * t1));
      }
      const C t = cadd(v1, v2);
      const C m1 = mk2<T>(v0.x - T(0.5) * t.x, v0.y - T(0.5) * t.y);
      C d = csub(v1, v2);
      const T s = T(0.86602540378443864676372317075294);
      d = mulI<DIR>(mk2<T>(s * d.x, s * d.y));
      dst[0] = cadd(v0, t);
      dst[Ns] = cadd(m1, d);
      dst[2 * Ns] = csub(m1, d);
    } else { // generic small radix (5, 7): direct DFT with roots taken from the twiddle table
      C v[7];
      for (int p = 0; p < R; p++) {
        v[p] = src[j + p * nb];
        if (Ns > 1 && p > 0) v[p] = cmul(v[p], twid(p * k * twStep));
      }
      const int rootStep = n / R;
      for (int q = 0; q < R; q++) {
        C acc = v[0];
        for (int p = 1; p < R; p++) acc = cadd(acc, cmul(v[p], twid(((p * q) % R) * rootStep)));
        dst[q * Ns] = acc;
      }
    }
  }
}

// All stages of `nf` transforms; data starts in buf0, returns the buffer that holds the result.
// Block-wide barriers inside: every thread of the CTA must call it.
template <class T, int DIR>
__device__ __forceinline__ typename Vec2<T>::type *fftInShared(typename Vec2<T>::type *buf0,
                                                               typename Vec2<T>::type *buf1, const FftAxis &ax,
                                                               int fstride, int nf,
                                                               const typename Vec2<T>::type *__restrict__ tw) {
  using C = typename Vec2<T>::type;
  C *a = buf0, *b = buf1;
  int Ns = 1;
  for (int s = 0; s < ax.nstages; s++) {
    const int R = ax.radix[s];
    stockhamStage<T, DIR>(a, b, ax.n, fstride, nf, R, Ns, tw);
    __syncthreads();
    Ns *= R;
    C *t = a; a = b; b = t;
  }
  return a;
}

} // namespace ub200
