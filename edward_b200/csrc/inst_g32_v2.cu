// Instantiates k_hmc<G=32, V=2, K> for every K tier (one translation unit per (G,V) so they build in parallel).
#include "chain.cuh"

namespace edhmc {
const void* lookup_g32_v2(int K) {
  switch (K) {
    case 1: return reinterpret_cast<const void*>(&k_hmc<32, 2, 1>);
    case 2: return reinterpret_cast<const void*>(&k_hmc<32, 2, 2>);
    case 4: return reinterpret_cast<const void*>(&k_hmc<32, 2, 4>);
    case 8: return reinterpret_cast<const void*>(&k_hmc<32, 2, 8>);
    case 16: return reinterpret_cast<const void*>(&k_hmc<32, 2, 16>);
    case 24: return reinterpret_cast<const void*>(&k_hmc<32, 2, 24>);
    case 27: return reinterpret_cast<const void*>(&k_hmc<32, 2, 27>);
    case 32: return reinterpret_cast<const void*>(&k_hmc<32, 2, 32>);
    default: return nullptr;
  }
}
}  // namespace edhmc
