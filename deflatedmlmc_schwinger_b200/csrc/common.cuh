// common.cuh -- scalar/complex helpers shared by all sm_100a kernels of libdmlmc_sm100.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmlmc {

// complex scalar of precision T, interleaved (re, im)
template <typename T> struct alignas(2 * sizeof(T)) Cx { T re, im; };

template <typename T> __host__ __device__ __forceinline__ Cx<T> cx(T re, T im) { Cx<T> r; r.re = re; r.im = im; return r; }
template <typename T> __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
  return cx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> __device__ __forceinline__ Cx<T> cconj(Cx<T> a) { return cx<T>(a.re, -a.im); }

// NC complex columns handled by one thread, loaded/stored as one 16-byte vector
// (c64: NC = 2 -> float4, c128: NC = 1 -> double2).  NC = 1 for c64 is the odd-k fallback.
template <typename T, int NC> struct alignas(2 * NC * sizeof(T)) Pack { T d[2 * NC]; };

template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> pzero() {
  Pack<T, NC> p;
#pragma unroll
  for (int i = 0; i < 2 * NC; ++i) p.d[i] = T(0);
  return p;
}
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> padd(Pack<T, NC> a, Pack<T, NC> b) {
#pragma unroll
  for (int i = 0; i < 2 * NC; ++i) a.d[i] += b.d[i];
  return a;
}
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> psub(Pack<T, NC> a, Pack<T, NC> b) {
#pragma unroll
  for (int i = 0; i < 2 * NC; ++i) a.d[i] -= b.d[i];
  return a;
}
// multiply by +i / -i (exact)
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> pmul_i(Pack<T, NC> a) {
  Pack<T, NC> r;
#pragma unroll
  for (int c = 0; c < NC; ++c) { r.d[2 * c] = -a.d[2 * c + 1]; r.d[2 * c + 1] = a.d[2 * c]; }
  return r;
}
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> pmul_mi(Pack<T, NC> a) {
  Pack<T, NC> r;
#pragma unroll
  for (int c = 0; c < NC; ++c) { r.d[2 * c] = a.d[2 * c + 1]; r.d[2 * c + 1] = -a.d[2 * c]; }
  return r;
}
// acc += s * x   (complex scalar s applied to every column of the pack): 4 FMA per column
template <typename T, int NC> __device__ __forceinline__ void pfma(Pack<T, NC>& acc, Cx<T> s, const Pack<T, NC>& x) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    T xr = x.d[2 * c], xi = x.d[2 * c + 1];
    acc.d[2 * c]     = fma(s.re, xr, fma(-s.im, xi, acc.d[2 * c]));
    acc.d[2 * c + 1] = fma(s.re, xi, fma(s.im, xr, acc.d[2 * c + 1]));
  }
}
// acc -= s * x
template <typename T, int NC> __device__ __forceinline__ void pfms(Pack<T, NC>& acc, Cx<T> s, const Pack<T, NC>& x) {
  pfma<T, NC>(acc, cx<T>(-s.re, -s.im), x);
}
// acc += conj(s) * x
template <typename T, int NC> __device__ __forceinline__ void pfma_conj(Pack<T, NC>& acc, Cx<T> s, const Pack<T, NC>& x) {
  pfma<T, NC>(acc, cx<T>(s.re, -s.im), x);
}
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> pscale(Cx<T> s, const Pack<T, NC>& x) {
  Pack<T, NC> r = pzero<T, NC>();
  pfma<T, NC>(r, s, x);
  return r;
}

template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> ldp(const Pack<T, NC>* p, size_t idx) { return p[idx]; }
template <typename T, int NC> __device__ __forceinline__ Pack<T, NC> ldp_ro(const Pack<T, NC>* p, size_t idx) {
  // read-only path for data never written by the same kernel
  Pack<T, NC> r;
  if constexpr (sizeof(Pack<T, NC>) == 16) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + idx));
    r = *reinterpret_cast<Pack<T, NC>*>(&v);
  } else if constexpr (sizeof(Pack<T, NC>) == 8) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p + idx));
    r = *reinterpret_cast<Pack<T, NC>*>(&v);
  } else {
    r = p[idx];
  }
  return r;
}
template <typename T> __device__ __forceinline__ Cx<T> ldc_ro(const Cx<T>* p, size_t idx) {
  Cx<T> r;
  if constexpr (sizeof(Cx<T>) == 16) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + idx));
    r = *reinterpret_cast<Cx<T>*>(&v);
  } else {
    float2 v = __ldg(reinterpret_cast<const float2*>(p + idx));
    r = *reinterpret_cast<Cx<T>*>(&v);
  }
  return r;
}

}  // namespace dmlmc
