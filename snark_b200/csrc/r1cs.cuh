// Device-resident R1CS matrices and Groth16 proving key (the handles behind b2s_r1cs / b2s_pk).
#pragma once
#include "common.cuh"

struct b2s_r1cs {
    uint64_t n_rows = 0, n_instance = 0, n_witness = 0;
    uint32_t log_domain = 0;          // domain = next_pow2(n_rows + n_instance)
    uint64_t nnz[3] = {0, 0, 0};
    // CSR per matrix; coefficients interned: id 0 == ONE (multiplication skipped, as
    // relations/src/sr1cs/mod.rs:42-46 does), other ids index `pool`.
    b2s::DevBuf row_ptr[3];           // uint64[n_rows + 1]
    b2s::DevBuf col[3];               // uint32[nnz]
    b2s::DevBuf coeff_id[3];          // uint32[nnz]
    b2s::DevBuf pool;                 // Fr[pool_size]
    uint32_t pool_size = 0;
};

struct b2s_pk {
    uint64_t n_instance = 0, n_witness = 0, domain_size = 0;
    b2s::DevBuf consts_g1;            // alpha_g1, beta_g1, delta_g1 (affine)
    b2s::DevBuf consts_g2;            // beta_g2, delta_g2 (affine)
    b2s::DevBuf a_query, b_g1_query, b_g2_query, h_query, l_query;
    // fixed-base window table of h_query (msm_precompute): [h_pre.nwin][h_len] affine points; empty when switched off / too big
    b2s::DevBuf h_table;
    b2s::MsmPre h_pre{0, 0, 0};
    uint32_t a_ext = 0, b1_ext = 0, b2_ext = 0;   // 2 when this shard owns the end of the range (extra delta pairs)
    uint64_t a_off = 0, a_len = 0, b1_off = 0, b1_len = 0, b2_off = 0, b2_len = 0, h_off = 0, h_len = 0, l_off = 0, l_len = 0;
};

namespace b2s {
int32_t r1cs_upload(Ctx* c, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness, const uint64_t* const row_ptr[3],
                    const uint32_t* const col[3], const void* const coeff[3], b2s_r1cs** out);
// lcmap.cu: the same handle from the constraint system's LcMap, CSR built by kernels
int32_t r1cs_upload_lcmap(Ctx* c, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness, const uint64_t* const args[3],
                          uint64_t n_lcs, const uint64_t* lc_offsets, const uint64_t* lc_vars, const uint32_t* lc_coeffs,
                          const void* pool, uint32_t pool_len, b2s_r1cs** out);
// out_k: device arrays with at least n_rows elements each
int32_t spmv_run(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* out_a, void* out_b, void* out_c);
// h_dev: device array of domain elements (output); z_dev: n_instance + n_witness elements
int32_t witness_map_run(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* h_dev);
int32_t pk_upload(Ctx* c, const b2s_pk_desc* d, int32_t mem, b2s_pk** out);
int32_t groth16_setup(Ctx* c, const b2s_r1cs* m, const void* trapdoor_host, b2s_pk** out_pk, void* o_alpha_g1, void* o_beta_g2,
                      void* o_gamma_g2, void* o_delta_g2, void* o_gamma_abc);
int32_t pk_query_download(Ctx* c, const b2s_pk* pk, int which, void* out_host, uint64_t cap_bytes);
// Where the h-query MSM of a shard takes its scalars from: the default computes the whole h on this GPU (replicated
// witness_map); the multi-GPU group computes it distributed and hands back this rank's coefficient slab (group.cu).
struct HSource {
    // z_dev: the full assignment on the device.  On return *h_for_shard points at the scalars matching pk->h_query
    // (pk->h_len elements starting at coefficient pk->h_off); the memory stays valid until the HSource is destroyed.
    virtual int32_t get(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_dev, const void** h_for_shard) = 0;
    virtual ~HSource() = default;
};
// z either as two host pieces (z_dev == nullptr) or as one device array; hs == nullptr: replicated witness_map
int32_t groth16_shard(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_inst_host, const void* z_wit_host,
                      const void* z_dev, const void* r_host, const void* s_host, void* g1_partials_dev /*4 xyzz*/,
                      void* g2_partial_dev /*1 xyzz*/, HSource* hs = nullptr);
int32_t groth16_finish(Ctx* c, const b2s_pk* pk, const void* g1_partials_dev, const void* g2_partials_dev, uint32_t n_shards,
                       const void* r_host, const void* s_host, void* out_a, void* out_b, void* out_c);
}  // namespace b2s
