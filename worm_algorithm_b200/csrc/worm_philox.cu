// worm_philox.cu -- chooses between the two production kernels (see worm_lat.cu, worm_wide.cu).
#include "worm_run.cuh"

namespace worm {

cudaError_t launch_philox(const PhiloxArgs &a, int lanes_per_chain, bool lowt, int smem_optin, int sm_count,
                          cudaStream_t st, const char **why)
{
    *why = "";
    int G = lanes_per_chain;
    if (G == 0) {
        // Latency mode while the batch fits one block per SM.  Preferred: 32 chains per walker with a
        // whole SM sub-partition to itself (code 2); else among 8 / 4 / 1 chains per walker (codes
        // 4 / 8 / 32) the one with the fewest chains per walker that still fits one wave.
        G = 1;
        for (int cand : {4, 8, 32, 2}) {
            const size_t need = lat_smem_bytes(a.L, cand, smem_optin);
            if (need == 0) continue;
            const int64_t bch = lat_chains_per_block(a.L, cand, smem_optin);
            const int64_t blocks = (a.n_chains + bch - 1) / bch;
            if (blocks <= sm_count && (cand != 2 || a.n_chains >= 16 * bch)) G = cand;
        }
    }
    if (G == 2 || G == 4 || G == 8 || G == 32) {
        if (a.L < 4 || (a.L & (a.L - 1))) { *why = "latency mode (lanes_per_chain 2/4/8/32) needs a power-of-two L >= 4"; return cudaErrorInvalidValue; }
        if (lat_smem_bytes(a.L, G, smem_optin) == 0) { *why = "lattice too large for latency mode with this lanes_per_chain"; return cudaErrorInvalidValue; }
        return launch_philox_lat(a, G, lowt, smem_optin, st);
    }
    if (G != 1 && G != -1) { *why = "lanes_per_chain must be 0, 1, 2, 4, 8, 32 or -1"; return cudaErrorInvalidValue; }
    return launch_philox_wide(a, G == -1, smem_optin, st);
}

}  // namespace worm
