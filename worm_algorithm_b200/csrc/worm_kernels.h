// worm_kernels.h -- launch prototypes between the C-ABI layer (c_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "worm_common.cuh"

namespace worm {

constexpr int MT_STRIDE = 8;    // int64 words per chain in the deterministic result record
constexpr int ACC_STRIDE = 8;   // int64 words per chain in the production accumulator record
constexpr int ACC_STATUS = 6;   // sticky per-chain status word of that record (0 = fine; see WORM_ACC_STATUS)

struct PhiloxArgs {
    int L;
    int algo;                 // 0 Taylor, 1 binary
    int64_t n_chains;
    uint64_t chain_id0;
    const PhiloxChain *prm;   // [n]
    uint64_t step_begin, step_end, therm;
    uint8_t *bonds;           // [n][2N] in/out
    int32_t *head, *tail;     // [n] in/out
    int64_t *acc;             // [n][ACC_STRIDE] in/out
    const int32_t *group_index;   // [n] or null
    int64_t *hist;            // [n_groups][N] or null
    int64_t *group_acc;       // [n_groups][ACC_STRIDE] or null
    uint64_t dump_every;
    int32_t max_dumps;
    int32_t *n_dumps;
    int64_t *events;
    uint8_t *dumps;
};

cudaError_t launch_mt(int L, int64_t n_chains, int64_t n_pad, const MtChain *d_prm,
                      uint64_t num_steps, uint64_t therm, int bond_flag,
                      uint8_t *d_bonds, int64_t *d_out,
                      int32_t max_events, int32_t *d_n_events, int64_t *d_events, uint8_t *d_dumps,
                      int32_t max_trace, int32_t *d_n_trace, int64_t *d_trace,
                      uint32_t *d_mt_state, int smem_optin, cudaStream_t st);

// lanes_per_chain: 0 = choose; 4 / 8 / 32 = latency mode with that many lanes per chain; 1 = thread
// per chain, bonds in shared memory when they fit; -1 = thread per chain, bonds in global memory;
// -2 = thread per chain, packed bonds in shared memory (worm_pack.cu).
// lowt: some chain has kfix > 2^32 or tfix < 2^32 (T < 1).  smem_optin = max opt-in dynamic shared
// memory per block.
cudaError_t launch_philox(const PhiloxArgs &a, int lanes_per_chain, bool lowt, int smem_optin, int sm_count,
                          cudaStream_t st, const char **why);

// the variant launch_philox runs for lanes_per_chain = 0 (one of 2, 4, 8, 32, 1, -2)
int choose_philox_variant(int L, int algo, int64_t n_chains, bool hist, int smem_optin, int sm_count);

// the two kernels behind launch_philox.  G = lanes_per_chain of the C ABI (4 / 8 / 32 <-> 8 / 4 / 1
// chains per walker warp of the latency kernel).
size_t lat_smem_bytes(int L, int G, int algo, bool hist, int smem_optin);   // 0 if this L / G cannot run in latency mode
int lat_chains_per_block(int L, int G, int algo, bool hist, int smem_optin);
// chains per block, walker groups per block and bond layout (0 dual, 1 byte per bond, 2 nibble) of a variant; false: cannot run
bool lat_info(int L, int G, int algo, bool hist, int smem_optin, int *chains_per_block, int *walkers, int *layout);
cudaError_t launch_philox_lat(const PhiloxArgs &a, int G, bool lowt, int smem_optin, cudaStream_t st);
cudaError_t launch_philox_wide(const PhiloxArgs &a, bool force_global, int smem_optin, cudaStream_t st);
// packed throughput kernel (worm_pack.cu): thread per chain, 4 bits (Taylor) / 1 bit (binary) per bond in shared memory
int pack_chains_per_block(int L, int algo, int64_t n_chains, int smem_optin, int sm_count);   // 0: this L cannot run packed
cudaError_t launch_philox_pack(const PhiloxArgs &a, bool lowt, int smem_optin, int sm_count, cudaStream_t st);

cudaError_t launch_image(int L, int64_t n, const uint8_t *d_bonds, int32_t *d_img, cudaStream_t st);
cudaError_t launch_block(int W, int64_t n, const int32_t *d_in, int32_t *d_out, int dbv, int variant,
                         cudaStream_t st);
cudaError_t launch_count(int W, int64_t n, const int32_t *d_img, int64_t *d_count, cudaStream_t st);
// dumps -> bond counts of `levels` blocking levels (+ optional int32 images per level) in one launch
cudaError_t launch_pipeline(int L, int64_t n, const uint8_t *d_bonds, int levels, int64_t *d_counts,
                            int32_t *const *h_level_images, int smem_optin, cudaStream_t st);
cudaError_t launch_boundaries(int Nx, int Ny, int64_t n, int m, const double *d_img, const double *d_cutoffs, int32_t *d_out,
                              cudaStream_t st);

// leading principal component + delete-one-block jackknife (pca.cu)
constexpr int PCA_MAX_BLOCKS = 64;
size_t pca_workspace_bytes(int d, int64_t n, int nb, bool with_iteration);
// stage timing, off by default: slots = {pack, host exactness gate + tensor map, Gram kernel, sum + power
// iteration + finish, unused}
constexpr int PCA_TIMING_SLOTS = 5;
void pca_timing(int enable);
void pca_last_timing(double *ms);
int pca_ldg(int d);   // leading dimension of the Gram matrices: d rounded up to a multiple of 4
cudaError_t launch_pca_gram(int d, int64_t n, const int32_t *d_img, int nb, int32_t *d_gram, void *d_ws, int sm_count,
                            cudaStream_t st, const char **why);
cudaError_t launch_pca(int d, int64_t n, const int32_t *d_img, int nb, int max_iters, double tol, double *d_explained,
                       double *d_sigma2, double *d_component, int *h_iters, double *h_residual, void *d_ws, int sm_count,
                       cudaStream_t st, const char **why);

}  // namespace worm
