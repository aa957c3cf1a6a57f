// Shared helpers for the PINN kernels.  Everything here compiles both as sm_100a
// device code (nvcc) and as plain host C++ (g++, -DPINN_EMU) so that the warp
// programs can be executed lane-by-lane on the CPU by tests/emu (test
// infrastructure only; the product never runs the host build).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/pinn_b200.h"

#if defined(__CUDACC__)
#define PINN_HD __host__ __device__ __forceinline__
#define PINN_D __device__ __forceinline__
#else
#define PINN_HD inline
#define PINN_D inline
#endif

namespace pinn {

struct alignas(16) f4 {
  float x, y, z, w;
};

PINN_HD f4 ld4(const float* p) {
#if defined(__CUDA_ARCH__)
  float4 v = *reinterpret_cast<const float4*>(p);
  return f4{v.x, v.y, v.z, v.w};
#else
  return f4{p[0], p[1], p[2], p[3]};
#endif
}
PINN_HD void st4(float* p, const f4& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float4*>(p) = make_float4(v.x, v.y, v.z, v.w);
#else
  p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
#endif
}
PINN_HD float& f4_at(f4& v, int i) { return (&v.x)[i]; }

// A pair of FP32 values that travels in one 64-bit register: jets with an even number of channels
// are stored as channel pairs so that the hot loops issue packed FMAs (fma.rn.f32x2 -> SASS FFMA2:
// two FMAs per lane per issue slot; a scalar operand is broadcast to both halves for free).
struct alignas(8) f2 {
  float x, y;
};

// c + a * s (s broadcast)
PINN_HD f2 fma2s(const f2& a, float s, const f2& c) {
#if defined(__CUDA_ARCH__)
  unsigned long long aa, ss, cc, rr;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(aa), "l"(ss), "l"(cc));
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rr));
  return r;
#else
  return f2{fmaf(a.x, s, c.x), fmaf(a.y, s, c.y)};
#endif
}
// c + a * b (elementwise)
PINN_HD f2 fma2(const f2& a, const f2& b, const f2& c) {
#if defined(__CUDA_ARCH__)
  unsigned long long aa, bb, cc, rr;
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rr) : "l"(aa), "l"(bb), "l"(cc));
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rr));
  return r;
#else
  return f2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
}

PINN_HD f2 ld2(const float* p) {
#if defined(__CUDA_ARCH__)
  float2 v = *reinterpret_cast<const float2*>(p);
  return f2{v.x, v.y};
#else
  return f2{p[0], p[1]};
#endif
}
PINN_HD void st2(float* p, const f2& v) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
#else
  p[0] = v.x; p[1] = v.y;
#endif
}

PINN_HD uint32_t f_bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  union { float f; uint32_t u; } c;
  c.f = x;
  return c.u;
#endif
}
PINN_HD float bits_f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
#endif
}

// FP32-grade products on a TF32 tensor pipe ("3xTF32"): x = hi + lo with hi = x rounded to TF32 (10 mantissa bits, round
// half away: +0x1000 then mask) and lo = x - hi exactly (|lo| <= 2^-11 |x|, either sign: no systematic bias);
// a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, the pipe reading the upper 19 bits of every operand.  What is dropped:
// the truncation of lo (<= 2^-21 |x| per operand) and lo*lo (<= 2^-22 |ab|).
PINN_HD void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#if defined(PINN_SPLIT_TRUNC)  // experiment: hi by truncation (the pipe ignores the low bits of x itself): 2 instructions
  hi = f_bits(x);
  lo = f_bits(x - bits_f(f_bits(x) & 0xffffe000u));
#else
  hi = (f_bits(x) + 0x1000u) & 0xffffe000u;
  lo = f_bits(x - bits_f(hi));
#endif
}
// the same product in plain arithmetic (host emulation of the tensor pipe: tests/emu)
PINN_HD float mul_tf32x3(float a, float b) {
  uint32_t ah, al, bh, bl;
  split_tf32(a, ah, al);
  split_tf32(b, bh, bl);
  const float alt = bits_f(al & 0xffffe000u), blt = bits_f(bl & 0xffffe000u);
  return (alt * bits_f(bh) + bits_f(ah) * blt) + bits_f(ah) * bits_f(bh);
}

#if defined(__CUDA_ARCH__)
// D += A (16x8, row-major) * B (8x8, column-major) on the warp-level tensor path (SASS HMMA.1688.F32.TF32).  Lane
// (g = lane / 4, t = lane % 4) holds a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]; b0 = B[t][g],
// b1 = B[t+4][g]; d0 = D[g][2t], d1 = D[g][2t+1], d2 = D[g+8][2t], d3 = D[g+8][2t+1].
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

PINN_HD float fast_exp(float x) {
#if defined(__CUDA_ARCH__)
  return __expf(x);
#else
  return expf(x);
#endif
}
PINN_HD float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;  // MUFU.RCP, <= 1 ulp; (__frcp_rn is a multi-instruction IEEE sequence with a branch)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

// sigma, sigma', sigma'', sigma''' of the activation at z
// (idrlnet/architecture/layer.py:68-125: tanh, swish = silu = z*sigmoid(z), sin).
template <int ACT>
PINN_HD void act_derivs(float z, float& s0, float& s1, float& s2, float& s3) {
  if (ACT == PINN_ACT_TANH) {
    const float a = tanhf(z);
    s0 = a;
    s1 = 1.0f - a * a;
    s2 = -2.0f * a * s1;
    s3 = -2.0f * s1 * (s1 - 2.0f * a * a);
  } else if (ACT == PINN_ACT_SILU) {
    const float s = fast_rcp(1.0f + fast_exp(-z));
    const float p = s * (1.0f - s);          // s'
    const float q = p * (1.0f - 2.0f * s);   // s''
    const float r = p * (1.0f - 6.0f * p);   // s'''
    s0 = z * s;
    s1 = s + z * p;
    s2 = 2.0f * p + z * q;
    s3 = 3.0f * q + z * r;
  } else {
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(z, &sn, &cs);
#else
    sn = sinf(z); cs = cosf(z);
#endif
    s0 = sn; s1 = cs; s2 = -sn; s3 = -cs;
  }
}

// sigma and its first FIVE derivatives at z: the higher-order jets (pure derivatives up to order 4 along one input:
// u_xxx, u_xxxx of examples/euler_beam/euler_beam.py:47-54) need sigma forward and sigma' in the reverse pass.
// tanh with a = tanh z, w = 1 - a^2: w, -2aw, -2w(1-3a^2), 8aw(2-3a^2), 8w(15a^4-15a^2+2);
// sigmoid with p = s(1-s): p, p(1-2s), p(1-6p), p(1-2s)(1-12p), p(1-30p+120p^2); silu = z*sigmoid: f^(n) = n s^(n-1) + z s^(n).
template <int ACT>
PINN_HD void act_derivs6(float z, float (&s)[6]) {
  if (ACT == PINN_ACT_TANH) {
    const float a = tanhf(z), w = 1.0f - a * a, a2 = a * a;
    s[0] = a;
    s[1] = w;
    s[2] = -2.0f * a * w;
    s[3] = -2.0f * w * (1.0f - 3.0f * a2);
    s[4] = 8.0f * a * w * (2.0f - 3.0f * a2);
    s[5] = 8.0f * w * (15.0f * a2 * a2 - 15.0f * a2 + 2.0f);
  } else if (ACT == PINN_ACT_SILU) {
    const float g = fast_rcp(1.0f + fast_exp(-z));
    const float p = g * (1.0f - g), m = 1.0f - 2.0f * g;
    const float g1 = p, g2 = p * m, g3 = p * (1.0f - 6.0f * p), g4 = p * m * (1.0f - 12.0f * p),
                g5 = p * (1.0f - 30.0f * p + 120.0f * p * p);
    s[0] = z * g;
    s[1] = g + z * g1;
    s[2] = 2.0f * g1 + z * g2;
    s[3] = 3.0f * g2 + z * g3;
    s[4] = 4.0f * g3 + z * g4;
    s[5] = 5.0f * g4 + z * g5;
  } else {
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(z, &sn, &cs);
#else
    sn = sinf(z); cs = cosf(z);
#endif
    s[0] = sn; s[1] = cs; s[2] = -sn; s[3] = -cs; s[4] = sn; s[5] = cs;
  }
}

// Same quantities with hardware approximations (MUFU.TANH, ex2/rcp, sin/cos approx): ~2^-11
// relative error, used only on the TF32 tensor-core path whose multiplicands carry 10 mantissa bits.
template <int ACT>
PINN_HD void act_derivs_fast(float z, float& s0, float& s1, float& s2, float& s3) {
#if defined(__CUDA_ARCH__)
  if (ACT == PINN_ACT_TANH) {
    float a;
    asm("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(z));
    s0 = a;
    s1 = 1.0f - a * a;
    s2 = -2.0f * a * s1;
    s3 = -2.0f * s1 * (s1 - 2.0f * a * a);
  } else if (ACT == PINN_ACT_SIN) {
    const float sn = __sinf(z), cs = __cosf(z);
    s0 = sn; s1 = cs; s2 = -sn; s3 = -cs;
  } else {
    act_derivs<ACT>(z, s0, s1, s2, s3);
  }
#else
  act_derivs<ACT>(z, s0, s1, s2, s3);
#endif
}

// Smallest S >= n (floats) with S % 4 == 0 and (S / 4) odd: per-lane rows of S
// floats accessed as float4 are bank-conflict free (8 lanes of a 128-bit phase
// land on 8 distinct 4-bank groups).
PINN_HD int row_stride(int n) {
  int s = (n + 3) & ~3;
  if (((s >> 2) & 1) == 0) s += 4;
  return s;
}

}  // namespace pinn
