// Drop-in replacement for the reference's src/query_batch.cuh:
// `result_handler = query_batch<location>` (database.hpp:183-189) forwarded to the libmcb200-backed class.
#ifndef MC_QUERY_BATCH_H_
#define MC_QUERY_BATCH_H_
#define MCB200_IN_REFERENCE_TREE
#include "mcb200_shim.hpp"
namespace mc { template <class L> using query_batch = mcb200::query_batch<L>; }
#endif
