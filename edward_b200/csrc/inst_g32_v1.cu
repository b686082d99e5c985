// Instantiates k_hmc<G=32, V=1, K> for every K tier (one translation unit per (G,V) so they build in parallel).
#include "chain.cuh"

namespace edhmc {
const void* lookup_g32_v1(int K) {
  switch (K) {
    case 1: return reinterpret_cast<const void*>(&k_hmc<32, 1, 1>);
    case 2: return reinterpret_cast<const void*>(&k_hmc<32, 1, 2>);
    case 4: return reinterpret_cast<const void*>(&k_hmc<32, 1, 4>);
    case 8: return reinterpret_cast<const void*>(&k_hmc<32, 1, 8>);
    case 16: return reinterpret_cast<const void*>(&k_hmc<32, 1, 16>);
    case 32: return reinterpret_cast<const void*>(&k_hmc<32, 1, 32>);
    case 64: return reinterpret_cast<const void*>(&k_hmc<32, 1, 64>);
    default: return nullptr;
  }
}
}  // namespace edhmc
