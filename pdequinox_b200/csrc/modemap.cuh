// Internal: the retained-mode index map shared by the channel-mix kernels (modemix.cu: CUDA cores, tc_modemix.cu: tcgen05).
#pragma once
#include "common.cuh"
#include "spectral_plan.h"

namespace pdeqx {

struct ModeMap {
  int D;
  int m[PDEQX_MAX_DIMS];   // the reference's mode counts (weight layout)
  int nd[PDEQX_MAX_DIMS];  // data layout: rows per leading axis (>= 2 m_a), modes on the last axis (>= m_{D-1}); the
                           // excess positions are padding that carries zeros (spectral_plan.h)
  int64_t mprod;  // prod m_a
};

// p -> (g, loc): corner index and offset inside the corner's (m_0..m_{D-1}) block; false for a padding position
__device__ __forceinline__ bool mode_to_weight(const ModeMap& mm, int64_t p, int& g, int64_t& loc) {
  const int D = mm.D;
  int64_t k = p % mm.nd[D - 1];
  int64_t rest = p / mm.nd[D - 1];
  bool valid = k < mm.m[D - 1];
  g = 0;
  loc = k;
  int64_t stride = mm.m[D - 1];
  for (int a = D - 2; a >= 0; --a) {
    int j = (int)(rest % mm.nd[a]);
    rest /= mm.nd[a];
    valid = valid && j < 2 * mm.m[a];
    int hi = j >= mm.m[a];
    g |= hi << a;
    loc += (int64_t)(j - hi * mm.m[a]) * stride;
    stride *= mm.m[a];
  }
  return valid;
}

static inline ModeMap make_map(const pdeqx_spectral_plan* p) {
  ModeMap mm;
  mm.D = p->desc.ndim;
  mm.mprod = 1;
  for (int a = 0; a < PDEQX_MAX_DIMS; ++a) {
    mm.m[a] = a < mm.D ? p->desc.modes[a] : 1;
    mm.nd[a] = a < mm.D - 1 ? p->n2[a] : (a == mm.D - 1 ? p->mlast : 1);
    if (a < mm.D) mm.mprod *= p->desc.modes[a];
  }
  return mm;
}

}  // namespace pdeqx
