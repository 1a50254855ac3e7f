/*
 * worm_b200.h -- C ABI of libworm_b200.so: the B200 (sm_100a) replacement for the worm-algorithm
 * hot path of saforem2/worm_algorithm.
 *
 * The reference has no FFI: its boundary is `subprocess.Popen` of one single-chain executable per
 * temperature plus text files (worm_simulation.py:107-129; input line worm_simulation.py:53-55 parsed
 * at worm_ising_2d.cpp:44; result line worm_ising_2d.cpp:189-191).  These entry points are what a
 * binding would call instead; INTEGRATION.md shows the ctypes stub for worm_simulation.py.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer, `h_` = host pointer, `stream` = a
 *     cudaStream_t passed as void* (NULL = default stream).
 *   - the caller allocates every buffer and keeps ownership; the library never frees caller memory.
 *   - every function returns WORM_OK (0) or a negative WORM_E_* code; worm_last_error() gives the
 *     text for the calling thread.  No entry point has global mutable state besides that
 *     thread-local message, so calls on different streams/threads are independent.
 *   - device entry points are stream-ordered and do not synchronise, EXCEPT that host parameter
 *     arrays (h_T, h_seed) are consumed before the call returns.
 *   - `*_host` entry points take host buffers only, allocate and free their own device memory, and
 *     synchronise before returning: they are the complete drop-in for a caller without CUDA code.
 *
 * Lattice conventions (worm_ising_2d.cpp:81-95,214-244): N = L*L sites, site = y*L + x, 2N bonds,
 * bond 2*site = +x bond, bond 2*site+1 = +y bond; directions 0:+x 1:+y 2:-x 3:-y.
 */
#ifndef WORM_B200_H
#define WORM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WORM_B200_ABI_VERSION 1

enum {
    WORM_OK = 0,
    WORM_E_ARG = -1,       /* invalid argument (message says which)                              */
    WORM_E_CUDA = -2,      /* CUDA runtime error (message carries cudaGetErrorString)            */
    WORM_E_NOGPU = -3,     /* no sm_100 device / no device at all: there is NO CPU fallback      */
    WORM_E_WORKSPACE = -4  /* workspace too small (see worm_*_workspace_bytes)                   */
};

enum { WORM_ALGO_TAYLOR = 0,   /* integer bond occupations, the reference's chain (worm_ising_2d.cpp:139-155) */
       WORM_ALGO_BINARY = 1 }; /* binary bonds, tanh(K) acceptance (north star; same parity-projected law)   */

/* per-chain accumulator record written by worm_philox_run: int64 acc[n_chains][WORM_ACC_STRIDE] */
enum { WORM_ACC_Z = 0,      /* closed iterations measured          (Z of worm_ising_2d.cpp:130)   */
       WORM_ACC_NB = 1,     /* sum of Nb over them                 (-Nb_tot of :132)               */
       WORM_ACC_NB2 = 2,    /* sum of Nb^2                                                           */
       WORM_ACC_N = 3,      /* sum of n = number of odd-parity bonds (count_bonds.py:110-117)        */
       WORM_ACC_N2 = 4,     /* sum of n^2                                                            */
       WORM_ACC_ITERS = 5,  /* measured loop iterations (steps with step >= therm)                   */
       WORM_ACC_STATUS = 6, /* sticky, 0 = fine.  Set only by the packed kernel (lanes_per_chain -2 / chosen
                               by 0), whose Taylor chain keeps 4 bits per bond: bit 0 = the chain's input
                               held an occupation > 15 (binary chain: > 1) and was left untouched; bit 1 =
                               an occupation reached 16 during the run (results of that chain are void;
                               rerun it with lanes_per_chain 1, which holds 0..255).  Never seen for
                               T >= 1: the weight of occupation n falls like K^n / n!                  */
       WORM_ACC_STRIDE = 8 };

/* per-chain result record written by worm_mt_run: int64 out[n_chains][WORM_MT_STRIDE] */
enum { WORM_MT_Z = 0, WORM_MT_NB_TOT = 1 /* negative running sum, like the reference's Nb_tot */,
       WORM_MT_STEPS = 2, WORM_MT_NB_FINAL = 3, WORM_MT_STATUS = 4 /* 0 ok, 1 occupation overflow (>254) */,
       WORM_MT_STRIDE = 8 };

int worm_abi_version(void);
/* copies the calling thread's last error text into buf (NUL-terminated); returns its length */
int worm_last_error(char *buf, size_t buflen);
/* number of visible sm_100 devices (0 if none); never fails */
int worm_device_count(void);
/* SM count and max opt-in shared memory per block of `device` */
int worm_device_info(int device, int *sm_count, int *smem_optin_bytes, int *clock_khz);

/* ---------------------------------------------------------------------------------------------
 * Deterministic mode: n_chains independent copies of the reference loop (worm_ising_2d.cpp:47-57,
 * 111-157), one WARP per chain (every lane walks the loop redundantly, the 32 lanes share the MT19937
 * refill and the dump copies), each chain consuming its own MT19937 stream seeded like the reference
 * (init_genrand((unsigned long)seed), worm_ising_2d.cpp:47 / mt19937ar.c:68-81).  Bit-exact: bond
 * dumps, Z, Nb_tot, step_num and Nb equal the reference executable's for the same (L, T, steps, seed).
 *
 *   h_T[n], h_seed[n]        temperature and (double) seed per chain -- the input.txt fields
 *   num_steps, bond_flag     as in input.txt; a chain stops at its first closed iteration with
 *                            step >= num_steps (worm_ising_2d.cpp:133-134)
 *   d_bonds  u8 [n][2N]      out: final occupations (what the trailing dump :160-167 writes)
 *   d_out    i64 [n][8]      out: WORM_MT_* record
 *   events:  at every point the reference appends an observables row (:113-119) the kernel stores
 *            d_events[c][e] = {step_num, Z, Nb_tot} (values before that iteration's kick) for
 *            e < max_events and, if bond_flag != 0 and d_dumps != NULL, the occupations in
 *            d_dumps[c][e][2N].  d_n_events[c] counts all events, stored or not.
 *   trace:   optional N_b trace: d_trace[c][k] = {step_num, Nb} at the k-th closed iteration,
 *            k < max_trace; d_n_trace[c] counts them all.  NULL / 0 to skip.
 *   d_workspace              worm_mt_workspace_bytes(n) bytes
 * ------------------------------------------------------------------------------------------- */
size_t worm_mt_workspace_bytes(int64_t n_chains);
int worm_mt_run(int L, int64_t n_chains, const double *h_T, const double *h_seed,
                uint64_t num_steps, int bond_flag,
                uint8_t *d_bonds, int64_t *d_out,
                int32_t max_events, int32_t *d_n_events, int64_t *d_events, uint8_t *d_dumps,
                int32_t max_trace, int32_t *d_n_trace, int64_t *d_trace,
                void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Production mode: counter-based Philox4x32-10 (Salmon et al., SC'11), random stream VERSION 2 --
 * the normative statement is oracle/worm_oracle.c:worm_oracle_philox_run, and worm_philox_words()
 * below returns the words so that a binder can check its reading of this paragraph:
 *   ONE Philox call serves TWO worm steps.  For step s of the chain with global id c and seed k:
 *     counter = { lo32(s >> 1), hi32(s >> 1), lo32(c), hi32(c) },  key = { lo32(k), hi32(k) },
 *     output (x, y, z, w);  an even step uses (a, b) = (x, y), an odd step (a, b) = (z, w).
 *     a                      acceptance draw, all 32 bits                        (worm_ising_2d.cpp:151)
 *     b >> 30                direction 0:+x 1:+y 2:-x 3:-y                       (:137)
 *     (b >> 29) & 1          0 = increment proposal, 1 = decrement proposal      (:141)
 *     ((b << 3) * N) >> 32   kick site, from the low 29 bits of b (32-bit shift, 64-bit product)  (:128)
 *   Acceptance in integers with kfix = ceil(2^32 K), tfix = ceil(2^32 / K), hfix = ceil(2^32 tanh K)
 *   (worm_thresholds):  Taylor increment  (nb + 1) * a < kfix  (and nb < 255), decrement  a < nb * tfix;
 *   binary chain: an occupied bond is always taken, an empty one iff a < hfix.
 * A chain's trajectory is a pure function of (seed, global chain id, T): it does not depend on launch
 * geometry, kernel variant, chunking or the number of GPUs.
 *
 * Runs steps [step_begin, step_begin + n_steps) of every chain, resuming from (d_bonds, d_head,
 * d_tail) and leaving the new state there.  Iterations with step >= therm_steps are measured into
 * d_acc (added to what is there; see WORM_ACC_*).  Unlike the reference thermalisation
 * really excludes the warm-up from the sums, and a run is exactly n_steps long.
 *
 *   algo                     WORM_ALGO_TAYLOR or WORM_ALGO_BINARY
 *   chain_id0                global id of chain 0 of this call (chain c uses id chain_id0 + c)
 *   h_T[n], h_seed[n]        per-chain temperature and 64-bit seed
 *   d_group_index i32 [n]    optional: group (e.g. temperature index) of every chain, < n_groups
 *   d_hist i64 [n_groups][N] optional G(r) (needs d_group_index): every measured iteration adds 1 to
 *                            d_hist[group][dy*L + dx], (dx,dy) = head - tail mod L, taken before the
 *                            kick, so bin 0 collects Z.  NULL to skip.
 *   d_group_acc i64 [n_groups][8]  optional (needs d_group_index): what this call added to each chain's
 *                            WORM_ACC_* record is also added (atomically) to its group's record, so a
 *                            caller that only wants per-temperature sums reads n_groups records back.
 *   dump_every, max_dumps    a measured closed iteration with step % dump_every == 0 stores
 *                            d_events[c][e] = {step, Z so far, sum Nb so far} and (d_dumps != NULL)
 *                            the occupations d_dumps[c][e][2N]; d_n_dumps[c] (in/out) counts them.
 *                            dump_every = 0 disables.
 *   lanes_per_chain          kernel variant (results are identical for all of them):
 *                            0 = choose (the latency kernel whenever L is a power of two that fits shared memory:
 *                            32 chains per walker if the lattice allows, else one chain per walker, or the variant
 *                            holding the most chains per block once the batch exceeds one wave);
 *                            2 / 4 / 8 / 32 = latency kernel with 32 / 8 / 4 / 1 chains per
 *                            walker warp (power-of-two L; 2 gives the walker an SM sub-partition to itself);
 *                            1 = one thread per chain, a byte per bond in shared memory;  -1 = one thread per
 *                            chain, bonds left in global memory (taken automatically when a lattice does not fit);
 *                            -2 = packed kernel: one thread per chain, 4 bits (Taylor) / 1 bit (binary) per bond in
 *                            shared memory (power-of-two L; see WORM_ACC_STATUS) -- what 0 chooses whenever the
 *                            batch has more chains than the latency kernel can keep in flight
 *   d_workspace              worm_philox_workspace_bytes(n) bytes
 * ------------------------------------------------------------------------------------------- */
size_t worm_philox_workspace_bytes(int64_t n_chains);
int worm_philox_run(int L, int algo, int64_t n_chains, uint64_t chain_id0,
                    const double *h_T, const uint64_t *h_seed,
                    uint64_t step_begin, uint64_t n_steps, uint64_t therm_steps,
                    uint8_t *d_bonds, int32_t *d_head, int32_t *d_tail, int64_t *d_acc,
                    const int32_t *d_group_index, int32_t n_groups, int64_t *d_hist, int64_t *d_group_acc,
                    uint64_t dump_every, int32_t max_dumps, int32_t *d_n_dumps,
                    int64_t *d_events, uint8_t *d_dumps,
                    int lanes_per_chain, void *d_workspace, size_t workspace_bytes, void *stream);

/* the kernel variant lanes_per_chain = 0 runs for this batch on the current device (2 / 4 / 8 / 32 / 1 / -2) */
int worm_philox_choice(int L, int algo, int64_t n_chains, int hist, int *lanes_per_chain);

/* The two words (a, b) of worm step `step` of chain `chain_id` under key `seed` -- the stream every
 * production kernel consumes (host arithmetic only, no GPU needed; for binders' known-answer tests).
 * Known answer, seed 1, chain 0: step 0 (0xe3e80670, 0xe50a0ebc), step 1 (0x95f222c0, 0xb615aa27),
 * step 2 (0xac08141b, 0xdfc5ccbe), step 3 (0x79c07a47, 0xa7f66093). */
int worm_philox_words(uint64_t seed, uint64_t chain_id, uint64_t step, uint32_t *a, uint32_t *b);

/* fixed-point acceptance constants used by the production kernels for temperature T:
 * kfix = ceil(2^32/T), tfix = ceil(2^32 * (1/(1/T))), hfix = ceil(2^32 * tanh(1/T)) */
int worm_thresholds(double T, uint64_t *kfix, uint64_t *tfix, uint64_t *hfix);

/* ---------------------------------------------------------------------------------------------
 * Post-processing of dumped configurations.
 *
 * worm_image   u8 occupations [n][2N] -> int32 image [n][2L][2L], first index = x pixel:
 *              odd-parity bonds and both their end sites set to 1 (bonds.py:199-239,241-319).
 * worm_block   one RG blocking step, int32 [n][W][W] -> [n][W/2][W/2] (W = 2L, even):
 *              variant 0 = block_configs._block_config / iterated_blocking._block_config
 *                          (block_configs.py:6-71; double_bonds_value 0 = "1+1=0" xor scheme)
 *              variant 1 = Bonds._block_config_T (bonds.py:333-410; double_bonds_value = block_val)
 * worm_count   n[c] = sum of pixels with (i + j) odd of int32 [n][W][W] (count_bonds.py:110-117).
 * ------------------------------------------------------------------------------------------- */
int worm_image(int L, int64_t n, const uint8_t *d_bonds, int32_t *d_img, void *stream);
int worm_block(int W, int64_t n, const int32_t *d_in, int32_t *d_out,
               int double_bonds_value, int variant, void *stream);
int worm_count(int W, int64_t n, const int32_t *d_img, int64_t *d_count, void *stream);

/* worm_pipeline  the whole chain bonds.py:199-319 -> iterated_blocking.py:6-59 (x k) -> count_bonds.py:107-123 in
 *              ONE launch, for the "1+1=0" blocking (double_bonds_value 0) of power-of-two lattices:
 *              d_bonds u8 [n][2N] -> d_counts i64 [n][levels], d_counts[c][k] = number of odd-parity bond pixels
 *              of configuration c after k blocking steps (k = 0: the image of the dump itself; lattice size
 *              L >> k >= 2).  Identical to worm_count(worm_block^k(worm_image(.))) -- the kernel works on the
 *              2 L^2 parity BITS of a configuration (a blocking step XORs the two bonds leaving each 2x2 block)
 *              and reads 2 L^2 bytes per configuration instead of building the 16 L^2-byte int32 image pyramid.
 *              h_level_images: NULL, or a HOST array of `levels` device pointers; a non-NULL entry k receives the
 *              int32 images [n][2(L>>k)][2(L>>k)] of level k (what worm_image / worm_block would have produced). */
int worm_pipeline(int L, int64_t n, const uint8_t *d_bonds, int levels, int64_t *d_counts,
                  int32_t *const *h_level_images, void *stream);

/* worm_boundaries  boundaries between the black and white regions of thresholded grey images
 *              (image_analysis/image.py:46-80, Image._cutoff + Image.get_boundaries), all images x all
 *              cutoffs in one launch: d_img f64 [n][Nx][Ny] (the class's `_image`, 1 - bytes/255),
 *              d_cutoffs f64 [m], d_links int32 [n][m][2Nx][2Ny] in the bond-image convention of
 *              worm_image (sites at even-even, links at odd-parity pixels), so worm_block / worm_count
 *              apply to it unchanged.  The comparison `pixel < cutoff` is done in f64 as the reference
 *              does. */
int worm_boundaries(int Nx, int Ny, int64_t n, const double *d_img, int m, const double *d_cutoffs,
                    int32_t *d_links, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Leading principal component of a stack of images, with the reference's delete-one-block jackknife
 * (pca.py:97-147: TruncatedSVD(n_components=1).fit(data).explained_variance_[0] of the un-centred
 * [n][d] matrix, i.e. the variance of the projection on the leading right singular vector;
 * utils.py:90-104 block_resampling = the num_blocks KFold training sets; utils.py:123-135).
 *
 * d_images      int32 [n][d], small integers (|v| <= 127 and max|v|^2 * n < 2^31: the tensor cores work
 *               on int8 with int32 accumulators, so the Gram matrices are exact; anything else
 *               returns WORM_E_ARG)
 * num_blocks    1 .. 64 (1 = no resampling); n >= num_blocks
 * d_explained   f64 [1 + num_blocks]: [0] the full data set, [1 + b] with block b deleted
 * d_sigma2      f64 [1 + num_blocks] or NULL: leading eigenvalue of X^T X (sigma_1^2) of each set
 * d_component   f64 [d] or NULL: components_[0] of the full set, sign such that it sums to > 0
 * h_iters, h_residual   host outputs (may be NULL): power iterations used, last max_j |dv_j|
 * Synchronises the stream (convergence is checked on the host every few iterations).
 *
 * worm_pca_gram is the tensor-core stage alone: d_gram int32 [num_blocks][ldg][ldg], ldg = (d + 3) & ~3,
 * G_b = X_b^T X_b, rows / columns >= d zero.  Workspace: worm_pca_workspace_bytes
 * with gram_only = 1.
 * worm_pca_timing(1) makes the following worm_pca / worm_pca_gram calls record CUDA events around
 * their stages (and worm_pca_gram synchronise); worm_pca_last_timing returns, for the calling thread's
 * last call, milliseconds of {pack, host gate, Gram kernel, iteration stage} (4 values).
 * ------------------------------------------------------------------------------------------- */
size_t worm_pca_workspace_bytes(int d, int64_t n, int num_blocks, int gram_only);
void worm_pca_timing(int enable);
void worm_pca_last_timing(double *ms4);
int worm_pca(int d, int64_t n, const int32_t *d_images, int num_blocks, int max_iters, double tol,
             double *d_explained, double *d_sigma2, double *d_component, int32_t *h_iters,
             double *h_residual, void *d_workspace, size_t workspace_bytes, void *stream);
int worm_pca_gram(int d, int64_t n, const int32_t *d_images, int num_blocks, int32_t *d_gram,
                  void *d_workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer drop-ins (allocate, copy in, run, copy out, free, synchronise).
 * worm_sim_host replaces the per-temperature Popen loop of worm_simulation.py:107-129:
 *   rng_mode 0 = deterministic MT19937 (h_seed read as double bits via h_seed_f64),
 *   rng_mode 1 = Philox (h_seed_u64).  h_out is i64 [n][8]: WORM_MT_* or WORM_ACC_* records.
 *   h_bonds (u8 [n][2N], may be NULL) receives the final occupations.
 * ------------------------------------------------------------------------------------------- */
int worm_sim_host(int L, int64_t n_chains, const double *h_T, const double *h_seed_f64,
                  const uint64_t *h_seed_u64, uint64_t num_steps, uint64_t therm_steps,
                  int rng_mode, int algo, int64_t *h_out, uint8_t *h_bonds);
int worm_image_host(int L, int64_t n, const uint8_t *h_bonds, int32_t *h_img);
int worm_block_host(int W, int64_t n, const int32_t *h_in, int32_t *h_out,
                    int double_bonds_value, int variant);
int worm_count_host(int W, int64_t n, const int32_t *h_img, int64_t *h_count);
/* h_counts i64 [n][levels]: worm_pipeline on host buffers (counts only) */
int worm_pipeline_host(int L, int64_t n, const uint8_t *h_bonds, int levels, int64_t *h_counts);
int worm_boundaries_host(int Nx, int Ny, int64_t n, const double *h_img, int m, const double *h_cutoffs,
                         int32_t *h_links);
/* h_explained f64 [1 + num_blocks]; h_component f64 [d] or NULL */
int worm_pca_host(int d, int64_t n, const int32_t *h_images, int num_blocks, int max_iters, double tol,
                  double *h_explained, double *h_component, int32_t *h_iters, double *h_residual);

#ifdef __cplusplus
}
#endif
#endif /* WORM_B200_H */
