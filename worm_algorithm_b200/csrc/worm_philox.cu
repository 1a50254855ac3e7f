// worm_philox.cu -- chooses between the two production kernels (see worm_lat.cu, worm_wide.cu).
#include "worm_run.cuh"

namespace worm {

int choose_philox_variant(int L, int algo, int64_t n_chains, bool hist, int smem_optin, int sm_count)
{
    // A chain is a sequence of dependent steps: throughput = (chains in flight) / (cycles per step), capped by the
    // issue slots.  Every variant that can run is rated by that estimate (cycles per step measured on B200,
    // DESIGN.md 3) and the best one is taken:
    //  * latency kernel (power-of-two L that fits shared memory; few cycles per step, few chains per SM):
    //    32 chains per walker with an SM sub-partition to itself ("solo", code 2; dual layout up to L = 32, nibble
    //    layout at L = 64), or 8 / 4 / 1 chains per walker (codes 4 / 8 / 32) in up to four groups per block;
    //  * packed kernel (code -2; thread per chain, 4 / 1 bits per bond): ~3x the cycles per step, but 7 times
    //    (Taylor, L = 32) to 24 times (binary) the chains per SM.
    // Thread-per-chain on bytes (code 1) is left for lattices that are not a power of two or do not fit.
    int G = 1;
    double best = 0.0;
    if (L >= 4 && !(L & (L - 1))) {
        for (int g : {2, 32, 8, 4}) {
            int cb = 0, nw = 0, lay = 0;
            if (!lat_info(L, g, algo, hist, smem_optin, &cb, &nw, &lay)) continue;
            if (g == 2 && n_chains < 64) continue;
            double cyc;
            // G(r): the solo walker counts in shared-memory bins where they fit (L <= 32, +14 cycles), else the
            // accountants replay the displacement (+34 next to a solo walker, a few cycles next to the groups)
            if (g == 2) cyc = (lay == 0 ? 57.0 : 74.0) + (hist ? (L <= 32 ? 14.0 : 34.0) : 0.0);
            else {
                const double base = (g == 4) ? 75.0 : (g == 8) ? 70.0 : 56.0;          // one group per block
                cyc = base + 9.0 * (nw - 1) + (lay == 2 ? 4.0 : 0.0) + (hist ? 8.0 : 0.0);
            }
            const double inflight = (double)(n_chains < (int64_t)cb * sm_count ? n_chains : (int64_t)cb * sm_count);
            const double rate = inflight / cyc;
            if (rate > best * 1.02) { best = rate; G = g; }
        }
        const int pcb = pack_chains_per_block(L, algo, n_chains, smem_optin, sm_count);
        if (pcb) {
            const double warps = (double)((pcb + 31) / 32);
            const double cyc_pack = (algo == ALGO_TAYLOR ? 176.0 + 13.0 * (warps - 1.0) : 130.0 + 21.0 * (warps - 1.0)) *
                                    (hist ? 1.35 : 1.0);
            const double inflight = (double)(n_chains < (int64_t)pcb * sm_count ? n_chains : (int64_t)pcb * sm_count);
            if (inflight / cyc_pack > 1.1 * best) G = -2;
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
        if (lat_smem_bytes(a.L, G, a.algo, hist, smem_optin) == 0) { *why = "lattice too large for latency mode with this lanes_per_chain"; return cudaErrorInvalidValue; }
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
