// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the box
// sampler's point / signed-distance maps; host+device so tests can check the
// device sampler bit-for-bit against a host evaluation.
#pragma once
#include "pinn_common.h"

namespace pinn {

PINN_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

// key = seed (64 bit), counter = (index lo, index hi, stream lo, stream hi)
PINN_HD void philox4x32_10(uint64_t seed, uint64_t stream_id, uint64_t index, uint32_t (&out)[4]) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)stream_id, c3 = (uint32_t)(stream_id >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    philox_mulhilo(0xD2511F53u, c0, h0, l0);
    philox_mulhilo(0xCD9E8D57u, c2, h1, l1);
    const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [lo, hi): 24 random bits -> u in [0,1)
PINN_HD float box_point(uint32_t r, float lo, float hi) {
  const float u = (float)(r >> 8) * (1.0f / 16777216.0f);
  return lo + u * (hi - lo);
}

// Signed distance of an axis-aligned box, positive inside
// (idrlnet/geo_utils/geo_obj.py:30-41 for 1 dim, :109-167 for 2 dims).
PINN_HD float box_sdf(const float* x, const float* lo, const float* hi, int dims) {
  float out2 = 0.f, in_max = -3.0e38f;
  for (int d = 0; d < dims; ++d) {
    const float c = lo[d] + 0.5f * (hi[d] - lo[d]);
    const float diff = fabsf(x[d] - c) - (hi[d] - c);
    const float pos = diff > 0.f ? diff : 0.f;
    out2 += pos * pos;
    in_max = diff > in_max ? diff : in_max;
  }
  const float inside = in_max < 0.f ? in_max : 0.f;
  return -(sqrtf(out2) + inside);
}

}  // namespace pinn
