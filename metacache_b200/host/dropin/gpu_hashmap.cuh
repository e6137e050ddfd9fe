// Drop-in replacement for the reference's src/gpu_hashmap.cuh: with -DGPU_MODE the
// reference selects `feature_store = gpu_hashmap<feature, location>` (database.hpp:183-189); this header
// forwards that name to the libmcb200-backed class, speaking the reference's own types.
#ifndef MC_GPU_HASHMAP_H_
#define MC_GPU_HASHMAP_H_
#define MCB200_IN_REFERENCE_TREE
#include "mcb200_shim.hpp"
namespace mc { template <class K, class V> using gpu_hashmap = mcb200::gpu_hashmap<K, V>; }
#endif
