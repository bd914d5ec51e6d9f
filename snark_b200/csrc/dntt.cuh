// Distributed witness_map (dntt.cu): interface shared with the group driver (group.cu) and the test entry (api.cu).
#pragma once
#include <algorithm>
#include <vector>

#include "ntt.cuh"
#include "r1cs.cuh"

namespace b2s {

// The one data-path collective of the witness map: every played rank r hands `send[r]` = G consecutive blocks (block d for
// rank d) and receives into `recv[r]` block s from rank s.  `flush` marks the last call of a batch (NCCL: close the group).
struct DistExchange {
    virtual int32_t all_to_all(Ctx* c, const void* const* send, void* const* recv, size_t block_bytes, bool flush) = 0;
    virtual ~DistExchange() = default;
};

// domain 2^log_n over 2^lg ranks: n even (N1 = N2 keeps one distribution down the whole chain) and lg <= the smaller radix
bool dist_supported(uint32_t log_n, uint32_t lg);

// h = witness_map(m, z) over 2^lg ranks.  `ranks`: the ranks this process plays; h_slab_out[r]: device buffer of N / G
// elements receiving coefficients [rank N/G, (rank+1) N/G) of h.  z_dev: the full assignment on this device.
int32_t witness_map_dist(Ctx* c, const b2s_r1cs* m, const void* z_dev, uint32_t lg, const std::vector<uint32_t>& ranks, DistExchange* xch,
                         void* const* h_slab_out);
// all 2^lg ranks on this GPU (exchange = device copies): h_dev receives the whole h; must equal witness_map_run bit for bit
int32_t witness_map_sim(Ctx* c, const b2s_r1cs* m, const void* z_dev, uint32_t lg, void* h_dev);

}  // namespace b2s
