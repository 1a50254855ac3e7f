// worm_philox.cu -- chooses between the two production kernels (see worm_lat.cu, worm_wide.cu).
#include "worm_run.cuh"

namespace worm {

int choose_philox_variant(int L, int algo, int64_t n_chains, bool hist, int smem_optin, int sm_count)
{
    struct { int L, algo; int64_t n_chains; } a{L, algo, n_chains};
    int G = 0;
    if (G == 0) {
        // A chain is a sequence of dependent steps: throughput = (chains in flight) / (cycles per step), capped by
        // the issue slots.  Two families, chosen by that estimate (cycle counts measured on B200):
        //  * latency kernel (power-of-two L that fits shared memory; few cycles per step, few chains per SM):
        //    32 chains per walker ("solo", code 2) when the lattice allows it and the batch has at least two
        //    blocks; else one chain per walker (code 32), or the variant with the most chains per block
        //    (codes 8 / 4) once the batch exceeds one chain per walker of a full wave;
        //  * packed kernel (code -2; thread per chain, 4 / 1 bits per bond): ~3x the cycles per step, but 7 times
        //    (Taylor, L = 32) to 24 times (binary) the chains per SM.
        // Thread-per-chain on bytes (code 1) is left for lattices that are not a power of two or do not fit.
        G = 1;
        double rate_lat = 0.0;
        if (a.L >= 4 && !(a.L & (a.L - 1))) {
            int cb = 0;
            double cyc = 0.0;
            if (lat_smem_bytes(a.L, 2, hist, smem_optin) && a.n_chains >= 64) {
                G = 2; cb = 32; cyc = hist ? 94.0 : 58.0;
            } else {
                const int c32 = lat_chains_per_block(a.L, 32, hist, smem_optin);
                int best = c32;
                if (c32) G = 32;
                if (!c32 || a.n_chains > (int64_t)c32 * sm_count) {
                    for (int g : {8, 4}) {
                        const int c = lat_chains_per_block(a.L, g, hist, smem_optin);
                        if (c > best) { best = c; G = g; }
                    }
                }
                cb = best;
                cyc = (best <= 1 ? 56.0 : best <= 4 ? 72.0 : 95.0) + (hist ? 60.0 : 0.0);
            }
            if (cb) {
                const double inflight = (double)(a.n_chains < (int64_t)cb * sm_count ? a.n_chains : (int64_t)cb * sm_count);
                rate_lat = inflight / cyc;
            }
            const int pcb = pack_chains_per_block(a.L, a.algo, a.n_chains, smem_optin, sm_count);
            if (pcb) {
                const double warps = (double)((pcb + 31) / 32);
                const double cyc_pack = (a.algo == ALGO_TAYLOR ? 176.0 + 13.0 * (warps - 1.0) : 130.0 + 21.0 * (warps - 1.0)) *
                                        (hist ? 1.35 : 1.0);
                const double inflight = (double)(a.n_chains < (int64_t)pcb * sm_count ? a.n_chains : (int64_t)pcb * sm_count);
                if (inflight / cyc_pack > 1.1 * rate_lat) G = -2;
            }
        }
    }
    return G;
}

cudaError_t launch_philox(const PhiloxArgs &a, int lanes_per_chain, bool lowt, int smem_optin, int sm_count,
                          cudaStream_t st, const char **why)
{
    *why = "";
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    int G = lanes_per_chain;
    if (G == 0) G = choose_philox_variant(a.L, a.algo, a.n_chains, hist, smem_optin, sm_count);
    if (G == 2 || G == 4 || G == 8 || G == 32) {
        if (a.L < 4 || (a.L & (a.L - 1))) { *why = "latency mode (lanes_per_chain 2/4/8/32) needs a power-of-two L >= 4"; return cudaErrorInvalidValue; }
        if (lat_smem_bytes(a.L, G, hist, smem_optin) == 0) { *why = "lattice too large for latency mode with this lanes_per_chain"; return cudaErrorInvalidValue; }
        return launch_philox_lat(a, G, lowt, smem_optin, st);
    }
    if (G == -2) {
        if (pack_chains_per_block(a.L, a.algo, a.n_chains, smem_optin, sm_count) == 0) { *why = "packed mode (lanes_per_chain -2) needs a power-of-two L >= 4 whose packed bonds fit shared memory"; return cudaErrorInvalidValue; }
        return launch_philox_pack(a, lowt, smem_optin, sm_count, st);
    }
    if (G != 1 && G != -1) { *why = "lanes_per_chain must be 0, 1, 2, 4, 8, 32, -1 or -2"; return cudaErrorInvalidValue; }
    return launch_philox_wide(a, G == -1, smem_optin, st);
}

}  // namespace worm
