// Instantiates k_hmc<G=1, V=1, K> for every K tier (one translation unit per (G,V) so they build in parallel).
#include "chain.cuh"

namespace edhmc {
const void* lookup_g1_v1(int K) {
  switch (K) {
    case 1: return reinterpret_cast<const void*>(&k_hmc<1, 1, 1>);
    case 2: return reinterpret_cast<const void*>(&k_hmc<1, 1, 2>);
    case 4: return reinterpret_cast<const void*>(&k_hmc<1, 1, 4>);
    case 8: return reinterpret_cast<const void*>(&k_hmc<1, 1, 8>);
    case 16: return reinterpret_cast<const void*>(&k_hmc<1, 1, 16>);
    case 32: return reinterpret_cast<const void*>(&k_hmc<1, 1, 32>);
    case 64: return reinterpret_cast<const void*>(&k_hmc<1, 1, 64>);
    default: return nullptr;
  }
}
}  // namespace edhmc
