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
