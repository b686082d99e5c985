// Instantiates k_hmc<G=1, V=4, K> for every K tier (one translation unit per (G,V) so they build in parallel).
#include "chain.cuh"

namespace edhmc {
const void* lookup_g1_v4(int K) {
  switch (K) {
    case 1: return reinterpret_cast<const void*>(&k_hmc<1, 4, 1>);
    case 2: return reinterpret_cast<const void*>(&k_hmc<1, 4, 2>);
    case 4: return reinterpret_cast<const void*>(&k_hmc<1, 4, 4>);
    case 8: return reinterpret_cast<const void*>(&k_hmc<1, 4, 8>);
    case 12: return reinterpret_cast<const void*>(&k_hmc<1, 4, 12>);
    case 16: return reinterpret_cast<const void*>(&k_hmc<1, 4, 16>);
    default: return nullptr;
  }
}
}  // namespace edhmc
