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

// n = prod radix; prefer 4, then 2, 3, 5, 7, 11. Returns false when n has other prime factors.
inline bool factorize(int n, FftAxis &ax) {
  ax.n = n;
  ax.nstages = 0;
  int m = n;
  while (m % 4 == 0 && ax.nstages < kMaxStages) { ax.radix[ax.nstages++] = 4; m /= 4; }
  const int primes[5] = {2, 3, 5, 7, 11}; // nextFFTWiseSize3D (utils/Grid.cuh:142-213) emits 2^a 3^b 5^c 7^d 11^e
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
    } else if (R == 5) { // 5-point butterfly with the two cosines / sines of 2 pi / 5 (Rader-free, 8 real products per part)
      C v0 = src[j], v1 = src[j + nb], v2 = src[j + 2 * nb], v3 = src[j + 3 * nb], v4 = src[j + 4 * nb];
      if (Ns > 1) {
        const int t1 = k * twStep;
        v1 = cmul(v1, twid(t1));
        v2 = cmul(v2, twid(2 * t1));
        v3 = cmul(v3, twid(3 * t1));
        v4 = cmul(v4, twid(4 * t1));
      }
      const T c1 = T(0.30901699437494742410229341718282), c2 = T(-0.80901699437494742410229341718282);
      const T s1 = T(0.95105651629515357211643933337938), s2 = T(0.58778525229247312916870595463907);
      const C t1 = cadd(v1, v4), t2 = cadd(v2, v3), t3 = csub(v1, v4), t4 = csub(v2, v3);
      const C a1 = mk2<T>(v0.x + c1 * t1.x + c2 * t2.x, v0.y + c1 * t1.y + c2 * t2.y);
      const C a2 = mk2<T>(v0.x + c2 * t1.x + c1 * t2.x, v0.y + c2 * t1.y + c1 * t2.y);
      const C b1 = mulI<DIR>(mk2<T>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
      const C b2 = mulI<DIR>(mk2<T>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
      dst[0] = cadd(v0, cadd(t1, t2));
      dst[Ns] = cadd(a1, b1);
      dst[2 * Ns] = cadd(a2, b2);
      dst[3 * Ns] = csub(a2, b2);
      dst[4 * Ns] = csub(a1, b1);
    } else { // prime radix 7 / 11: the inputs p and R - p enter through their sum and difference, so that the outputs q and
             // R - q share (R - 1) / 2 real-by-complex products each instead of R - 1 complex products
      C v0 = src[j];
      C tp[5], tm[5];
      const int H = (R - 1) >> 1;
      C sum = v0;
      for (int p = 1; p <= H; p++) {
        C x = src[j + p * nb], y = src[j + (R - p) * nb];
        if (Ns > 1) {
          x = cmul(x, twid(p * k * twStep));
          y = cmul(y, twid((R - p) * k * twStep));
        }
        tp[p - 1] = cadd(x, y);
        tm[p - 1] = csub(x, y);
        sum = cadd(sum, tp[p - 1]);
      }
      dst[0] = sum;
      const int rootStep = n / R;
      for (int q = 1; q <= H; q++) {
        C a = v0, bsum = mk2<T>(T(0), T(0));
        for (int p = 1; p <= H; p++) {
          const C w = tw[((p * q) % R) * rootStep]; // (cos, -sin)(2 pi p q / R)
          a = mk2<T>(a.x + w.x * tp[p - 1].x, a.y + w.x * tp[p - 1].y);
          bsum = mk2<T>(bsum.x - w.y * tm[p - 1].x, bsum.y - w.y * tm[p - 1].y);
        }
        const C bi = mulI<DIR>(bsum);
        dst[q * Ns] = cadd(a, bi);
        dst[(R - q) * Ns] = csub(a, bi);
      }
    }
  }
}

// ---- compile-time specialised stages (power-of-two axes) ----
// Same Stockham indexing as stockhamStage, but N, the radix and Ns are template constants: every division is a
// shift/mask, the butterflies are fully unrolled, and an odd log2(N) is absorbed by ONE leading radix-8 stage
// (Ns = 1, so it needs no twiddles): 128 = 8*4*4 is three trips through shared memory instead of four.
template <int DIR, class C> __device__ __forceinline__ void radix4(C &v0, C &v1, C &v2, C &v3) {
  const C s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = mulI<DIR>(csub(v1, v3));
  v0 = cadd(s02, s13); v1 = cadd(d02, d13); v2 = csub(s02, s13); v3 = csub(d02, d13);
}
// multiply by exp(DIR * i * pi / 4) and exp(DIR * 3 i pi / 4)
template <int DIR, class T, class C> __device__ __forceinline__ C mulW8(C a) {
  const T h = T(0.70710678118654752440084436210485);
  return DIR < 0 ? mk2<T>(h * (a.x + a.y), h * (a.y - a.x)) : mk2<T>(h * (a.x - a.y), h * (a.x + a.y));
}
template <int DIR, class T, class C> __device__ __forceinline__ C mulW8c(C a) {
  const T h = T(0.70710678118654752440084436210485);
  return DIR < 0 ? mk2<T>(h * (a.y - a.x), -h * (a.x + a.y)) : mk2<T>(-h * (a.x + a.y), h * (a.x - a.y));
}

constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n / 2); }
constexpr bool fftFixedSupported(int n) { return n >= 8 && n <= 2048 && (n & (n - 1)) == 0; }

template <class T, int DIR, int N, int R, int Ns>
__device__ __forceinline__ void stockhamStageFixed(const typename Vec2<T>::type *__restrict__ a,
                                                   typename Vec2<T>::type *__restrict__ b, int fstride, int nf,
                                                   const typename Vec2<T>::type *__restrict__ tw) {
  using C = typename Vec2<T>::type;
  constexpr int nb = N / R, lnb = ilog2c(nb);
  constexpr int twStep = N / (Ns * R);
  auto twid = [&](int idx) {
    C t = __ldg(tw + idx);
    if (DIR > 0) t.y = -t.y;
    return t;
  };
  // work item w -> (transform f FASTEST, butterfly j): consecutive lanes touch consecutive transforms, i.e. shared
  // memory addresses fstride (odd) complex numbers apart -> distinct banks. With j fastest the Ns = 1 stage would
  // write R elements apart (8 * 16 B = every lane of a quarter warp in the same bank: measured 47 % replays).
  const int total = nf << lnb;
  const float invNf = 1.0f / (float)nf;
  for (int w = threadIdx.x; w < total; w += blockDim.x) {
    const int j = __float2int_rz(((float)w + 0.5f) * invNf), f = w - j * nf;
    const int k = j & (Ns - 1);
    const C *src = a + f * fstride + j;
    C *dst = b + f * fstride + (j - k) * R + k;
    if (R == 4) {
      C v0 = src[0], v1 = src[nb], v2 = src[2 * nb], v3 = src[3 * nb];
      if (Ns > 1) {
        const int t1 = k * twStep;
        v1 = cmul(v1, twid(t1)); v2 = cmul(v2, twid(2 * t1)); v3 = cmul(v3, twid(3 * t1));
      }
      radix4<DIR>(v0, v1, v2, v3);
      dst[0] = v0; dst[Ns] = v1; dst[2 * Ns] = v2; dst[3 * Ns] = v3;
    } else if (R == 8) {
      C v0 = src[0], v1 = src[nb], v2 = src[2 * nb], v3 = src[3 * nb];
      C v4 = src[4 * nb], v5 = src[5 * nb], v6 = src[6 * nb], v7 = src[7 * nb];
      if (Ns > 1) {
        const int t1 = k * twStep;
        v1 = cmul(v1, twid(t1)); v2 = cmul(v2, twid(2 * t1)); v3 = cmul(v3, twid(3 * t1)); v4 = cmul(v4, twid(4 * t1));
        v5 = cmul(v5, twid(5 * t1)); v6 = cmul(v6, twid(6 * t1)); v7 = cmul(v7, twid(7 * t1));
      }
      radix4<DIR>(v0, v2, v4, v6); // even half: E0..E3 in v0, v2, v4, v6
      radix4<DIR>(v1, v3, v5, v7); // odd half:  O0..O3 in v1, v3, v5, v7
      const C o1 = mulW8<DIR, T>(v3), o2 = mulI<DIR>(v5), o3 = mulW8c<DIR, T>(v7);
      dst[0] = cadd(v0, v1); dst[4 * Ns] = csub(v0, v1);
      dst[Ns] = cadd(v2, o1); dst[5 * Ns] = csub(v2, o1);
      dst[2 * Ns] = cadd(v4, o2); dst[6 * Ns] = csub(v4, o2);
      dst[3 * Ns] = cadd(v6, o3); dst[7 * Ns] = csub(v6, o3);
    } else { // R == 2
      C v0 = src[0], v1 = src[nb];
      if (Ns > 1) v1 = cmul(v1, twid(k * twStep));
      dst[0] = cadd(v0, v1); dst[Ns] = csub(v0, v1);
    }
  }
}

template <class T, int DIR, int N, int Ns>
__device__ __forceinline__ typename Vec2<T>::type *fftFixedFrom(typename Vec2<T>::type *a, typename Vec2<T>::type *b,
                                                                int fstride, int nf,
                                                                const typename Vec2<T>::type *__restrict__ tw) {
  if constexpr (Ns >= N) {
    return a;
  } else {
    constexpr int R = (Ns == 1 && (ilog2c(N) & 1)) ? 8 : 4;
    stockhamStageFixed<T, DIR, N, R, Ns>(a, b, fstride, nf, tw);
    __syncthreads();
    return fftFixedFrom<T, DIR, N, Ns * R>(b, a, fstride, nf, tw);
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

// NFIX > 0: axis length known at compile time (power of two) -> specialised stages; 0 -> generic
template <class T, int DIR, int NFIX>
__device__ __forceinline__ typename Vec2<T>::type *fftShared(typename Vec2<T>::type *buf0, typename Vec2<T>::type *buf1,
                                                             const FftAxis &ax, int fstride, int nf,
                                                             const typename Vec2<T>::type *__restrict__ tw) {
  if constexpr (NFIX > 0) return fftFixedFrom<T, DIR, NFIX, 1>(buf0, buf1, fstride, nf, tw);
  else return fftInShared<T, DIR>(buf0, buf1, ax, fstride, nf, tw);
}

} // namespace ub200
