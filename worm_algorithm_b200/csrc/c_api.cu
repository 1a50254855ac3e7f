// c_api.cu -- the extern "C" boundary declared in include/worm_b200.h.
//
// Host-side work done here and nowhere else: argument checks, the per-chain parameter records
// (temperature -> the doubles / fixed-point constants the kernels consume, computed exactly as the
// reference computes K and 1/K at worm_ising_2d.cpp:49-50), upload of those records into the
// caller's workspace, and kernel launches.  There is no CPU implementation behind these entry
// points: without an sm_100 device they return WORM_E_NOGPU.
#include "../../include/worm_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "worm_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(WORM_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DevInfo {
    int ok = 0, sm = 0, smem_optin = 0, clock_khz = 0, major = 0, minor = 0;
};

// Device attributes are queried once per device and cached: cudaDevAttrClockRate alone costs
// milliseconds per query, which would dwarf a 78 us blocking launch.
int current_device_info(DevInfo &d)
{
    constexpr int MAX_DEV = 64;
    static DevInfo cache[MAX_DEV];
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(WORM_E_NOGPU, "no CUDA device visible (%s); libworm_b200 has no CPU fallback", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= MAX_DEV) return fail(WORM_E_NOGPU, "device ordinal %d unsupported", dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!cache[dev].ok) {
            DevInfo q;
            cudaDeviceGetAttribute(&q.major, cudaDevAttrComputeCapabilityMajor, dev);
            cudaDeviceGetAttribute(&q.minor, cudaDevAttrComputeCapabilityMinor, dev);
            cudaDeviceGetAttribute(&q.sm, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&q.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            cudaDeviceGetAttribute(&q.clock_khz, cudaDevAttrClockRate, dev);
            if (q.major != 10)
                return fail(WORM_E_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", dev, q.major, q.minor);
            q.ok = 1;
            cache[dev] = q;
        }
        d = cache[dev];
    }
    return WORM_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void thresholds(double T, uint64_t &kfix, uint64_t &tfix, uint64_t &hfix)
{
    const double K = 1.0 / T;                 // worm_ising_2d.cpp:49
    const double inv_K = 1.0 / K;             // :50
    kfix = (uint64_t)std::ceil(std::ldexp(K, 32));
    tfix = (uint64_t)std::ceil(std::ldexp(inv_K, 32));
    hfix = (uint64_t)std::ceil(std::ldexp(std::tanh(K), 32));
}

bool bad_T(double T) { return !(T > 0.0) || !std::isfinite(T) || T < 1e-3 || T > 1e6; }

}  // namespace

extern "C" {

int worm_abi_version(void) { return WORM_B200_ABI_VERSION; }

int worm_last_error(char *buf, size_t buflen)
{
    const size_t n = strlen(g_err);
    if (buf && buflen) {
        const size_t k = n < buflen - 1 ? n : buflen - 1;
        memcpy(buf, g_err, k);
        buf[k] = 0;
    }
    return (int)n;
}

int worm_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < count; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int worm_device_info(int device, int *sm_count, int *smem_optin_bytes, int *clock_khz)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return fail(WORM_E_NOGPU, "device %d not available", device);
    }
    if (sm_count) cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, device);
    if (smem_optin_bytes) cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (clock_khz) cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, device);
    return WORM_OK;
}

int worm_thresholds(double T, uint64_t *kfix, uint64_t *tfix, uint64_t *hfix)
{
    if (bad_T(T)) return fail(WORM_E_ARG, "temperature %g out of range", T);
    uint64_t k, t, h;
    thresholds(T, k, t, h);
    if (kfix) *kfix = k;
    if (tfix) *tfix = t;
    if (hfix) *hfix = h;
    return WORM_OK;
}

// Philox4x32-10 on the host: the words of one worm step of the production stream (version 2, see the header)
int worm_philox_words(uint64_t seed, uint64_t chain_id, uint64_t step, uint32_t *a, uint32_t *b)
{
    const uint64_t pair = step >> 1;
    uint32_t c0 = (uint32_t)pair, c1 = (uint32_t)(pair >> 32), c2 = (uint32_t)chain_id, c3 = (uint32_t)(chain_id >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    if (a) *a = (step & 1) ? c2 : c0;
    if (b) *b = (step & 1) ? c3 : c1;
    return WORM_OK;
}

/* ------------------------------------------------------------------------------------------- */
size_t worm_mt_workspace_bytes(int64_t n_chains)
{
    if (n_chains <= 0) return 0;
    const size_t n_pad = align_up((size_t)n_chains, 32);
    return align_up(n_pad * sizeof(worm::MtChain), 256) + 256;     // chain parameter records (the MT19937 state lives in shared memory)
}

int worm_mt_run(int L, int64_t n_chains, const double *h_T, const double *h_seed,
                uint64_t num_steps, int bond_flag,
                uint8_t *d_bonds, int64_t *d_out,
                int32_t max_events, int32_t *d_n_events, int64_t *d_events, uint8_t *d_dumps,
                int32_t max_trace, int32_t *d_n_trace, int64_t *d_trace,
                void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (L < 2 || L > 1024) return fail(WORM_E_ARG, "L = %d out of range [2, 1024]", L);
    if (n_chains <= 0) return fail(WORM_E_ARG, "n_chains = %lld must be positive", (long long)n_chains);
    if (!h_T || !h_seed || !d_bonds || !d_out) return fail(WORM_E_ARG, "null required pointer");
    if (max_events < 0 || max_trace < 0) return fail(WORM_E_ARG, "negative buffer capacity");
    if (max_events > 0 && !d_n_events) return fail(WORM_E_ARG, "d_n_events required when max_events > 0");
    if (workspace_bytes < worm_mt_workspace_bytes(n_chains) || !d_workspace)
        return fail(WORM_E_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, worm_mt_workspace_bytes(n_chains));
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const size_t n_pad = align_up((size_t)n_chains, 32);
    std::vector<worm::MtChain> prm((size_t)n_chains);
    for (int64_t c = 0; c < n_chains; ++c) {
        const double T = h_T[c];
        if (bad_T(T)) return fail(WORM_E_ARG, "temperature[%lld] = %g out of range", (long long)c, T);
        if (!(h_seed[c] >= 0.0) || h_seed[c] >= 18446744073709551616.0)
            return fail(WORM_E_ARG, "seed[%lld] = %g must be in [0, 2^64)", (long long)c, h_seed[c]);
        worm::MtChain &p = prm[(size_t)c];
        p.T = T;
        p.K = 1.0 / T;                                         // worm_ising_2d.cpp:49
        p.inv_K = 1.0 / p.K;                                   // :50
        p.seed = (uint32_t)(unsigned long)h_seed[c];           // :47 -> mt19937ar.c:70 (low 32 bits)
        p.pad = 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)d_workspace;
    worm::MtChain *d_prm = (worm::MtChain *)ws;
    uint32_t *d_state = (uint32_t *)(ws + align_up(n_pad * sizeof(worm::MtChain), 256));
    cudaError_t e = cudaMemcpyAsync(d_prm, prm.data(), prm.size() * sizeof(worm::MtChain), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "upload chain parameters");
    const uint64_t therm = (uint64_t)(0.1 * (double)num_steps);   // worm_ising_2d.cpp:55
    e = worm::launch_mt(L, n_chains, (int64_t)n_pad, d_prm, num_steps, therm, bond_flag, d_bonds, d_out,
                        max_events, d_n_events, d_events, d_dumps, max_trace, d_n_trace, d_trace, d_state, dev.smem_optin, st);
    if (e != cudaSuccess) return cuda_fail(e, "worm_mt_kernel launch");
    // cudaMemcpyAsync from pageable memory has already staged `prm`; nothing else to wait for
    return WORM_OK;
}

/* ------------------------------------------------------------------------------------------- */
size_t worm_philox_workspace_bytes(int64_t n_chains)
{
    if (n_chains <= 0) return 0;
    return align_up((size_t)n_chains * sizeof(worm::PhiloxChain), 256) + 256;
}

int worm_philox_run(int L, int algo, int64_t n_chains, uint64_t chain_id0,
                    const double *h_T, const uint64_t *h_seed,
                    uint64_t step_begin, uint64_t n_steps, uint64_t therm_steps,
                    uint8_t *d_bonds, int32_t *d_head, int32_t *d_tail, int64_t *d_acc,
                    const int32_t *d_group_index, int32_t n_groups, int64_t *d_hist, int64_t *d_group_acc,
                    uint64_t dump_every, int32_t max_dumps, int32_t *d_n_dumps,
                    int64_t *d_events, uint8_t *d_dumps,
                    int lanes_per_chain, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (L < 3 || L > 1024) return fail(WORM_E_ARG, "L = %d out of range [3, 1024] (production mode)", L);
    if (algo != WORM_ALGO_TAYLOR && algo != WORM_ALGO_BINARY) return fail(WORM_E_ARG, "algo = %d", algo);
    if (n_chains <= 0) return fail(WORM_E_ARG, "n_chains = %lld must be positive", (long long)n_chains);
    if (!h_T || !h_seed || !d_bonds || !d_head || !d_tail || !d_acc) return fail(WORM_E_ARG, "null required pointer");
    if ((d_hist || d_group_acc) && !d_group_index) return fail(WORM_E_ARG, "d_hist / d_group_acc need d_group_index");
    if (d_group_index && n_groups <= 0) return fail(WORM_E_ARG, "n_groups must be positive with d_group_index");
    if (dump_every && (max_dumps <= 0 || !d_n_dumps)) return fail(WORM_E_ARG, "dump_every needs max_dumps > 0 and d_n_dumps");
    if (lanes_per_chain != 0 && lanes_per_chain != 1 && lanes_per_chain != 2 && lanes_per_chain != 4 &&
        lanes_per_chain != 8 && lanes_per_chain != 32 && lanes_per_chain != -1 && lanes_per_chain != -2)
        return fail(WORM_E_ARG, "lanes_per_chain = %d (0, 1, 2, 4, 8, 32, -1 or -2)", lanes_per_chain);
    if (step_begin + n_steps < step_begin) return fail(WORM_E_ARG, "step range overflows");
    if (workspace_bytes < worm_philox_workspace_bytes(n_chains) || !d_workspace)
        return fail(WORM_E_WORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, worm_philox_workspace_bytes(n_chains));
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    std::vector<worm::PhiloxChain> prm((size_t)n_chains);
    // a sweep has few distinct temperatures, in any order: the fixed-point constants (ceil / ldexp / tanh)
    // are computed once per temperature, looked up by the bits of T in a small open-addressed table (linear
    // probing: a direct-mapped one let two temperatures of an interleaved sweep evict each other chain after
    // chain -- 64 temperatures collide somewhere in 1024 slots more often than not -- which cost 30 ms of host
    // time per launch on a 10^6-chain batch); when it fills up, the rest is computed directly
    struct Slot { uint64_t bits; uint64_t k, t, h; bool used; };
    constexpr size_t NSLOT = 1024, MAXFILL = 768;
    std::vector<Slot> slots(NSLOT, Slot{0, 0, 0, 0, false});
    size_t filled = 0;
    bool lowt = false;            // some chain outside the T >= 1 fast path of the kernels
    for (int64_t c = 0; c < n_chains; ++c) {
        const double T = h_T[c];
        uint64_t bits;
        memcpy(&bits, &T, sizeof bits);
        size_t i = (size_t)((bits * 0x9E3779B97F4A7C15ull) >> 54);
        while (slots[i].used && slots[i].bits != bits) i = (i + 1) & (NSLOT - 1);
        Slot tmp{bits, 0, 0, 0, true};
        Slot *sl = &slots[i];
        if (!sl->used) {
            if (bad_T(T)) return fail(WORM_E_ARG, "temperature[%lld] = %g out of range", (long long)c, T);
            if (filled >= MAXFILL) sl = &tmp;                    // table full: do not insert
            else ++filled;
            thresholds(T, sl->k, sl->t, sl->h);
            sl->bits = bits;
            sl->used = true;
            if (sl->k > 0x100000000ull || sl->t < 0x100000000ull) lowt = true;
        }
        prm[(size_t)c] = worm::PhiloxChain{sl->k, sl->t, sl->h, h_seed[c]};
    }
    cudaStream_t st = (cudaStream_t)stream;
    worm::PhiloxChain *d_prm = (worm::PhiloxChain *)d_workspace;
    cudaError_t e = cudaMemcpyAsync(d_prm, prm.data(), prm.size() * sizeof(worm::PhiloxChain), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "upload chain parameters");
    worm::PhiloxArgs a;
    a.L = L; a.algo = algo; a.n_chains = n_chains; a.chain_id0 = chain_id0; a.prm = d_prm;
    a.therm = therm_steps;
    a.bonds = d_bonds; a.head = d_head; a.tail = d_tail; a.acc = d_acc;
    a.group_index = d_group_index; a.hist = d_hist; a.group_acc = d_group_acc;
    a.dump_every = dump_every; a.max_dumps = max_dumps; a.n_dumps = d_n_dumps; a.events = d_events; a.dumps = d_dumps;
    // One launch per chunk of <= 2^31 steps (round counters inside the kernels are 32-bit).  Chunking does
    // not change any result (counter-based stream).
    const uint64_t chunk = 0x80000000ull;
    uint64_t s = step_begin;
    const uint64_t s_end = step_begin + n_steps;
    do {
        a.step_begin = s;
        a.step_end = (s_end - s > chunk) ? s + chunk : s_end;
        const char *why = "";
        e = worm::launch_philox(a, lanes_per_chain, lowt, dev.smem_optin, dev.sm, st, &why);
        if (e != cudaSuccess) {
            if (why[0]) return fail(WORM_E_ARG, "%s", why);
            return cuda_fail(e, "worm kernel launch");
        }
        s = a.step_end;
    } while (s < s_end);
    // cudaMemcpyAsync from pageable memory has already staged `prm`; nothing else to wait for
    return WORM_OK;
}

int worm_philox_choice(int L, int algo, int64_t n_chains, int hist, int *lanes_per_chain)
{
    if (!lanes_per_chain) return fail(WORM_E_ARG, "null pointer");
    if (L < 3 || L > 1024 || n_chains <= 0) return fail(WORM_E_ARG, "L = %d / n_chains = %lld out of range", L, (long long)n_chains);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    *lanes_per_chain = worm::choose_philox_variant(L, algo, n_chains, hist != 0, dev.smem_optin, dev.sm);
    return WORM_OK;
}

/* ------------------------------------------------------------------------------------------- */
int worm_image(int L, int64_t n, const uint8_t *d_bonds, int32_t *d_img, void *stream)
{
    if (L < 3 || L > 256) return fail(WORM_E_ARG, "L = %d out of range [3, 256] for imaging", L);
    if (n < 0 || (n > 0 && (!d_bonds || !d_img))) return fail(WORM_E_ARG, "bad n / null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    cudaError_t e = worm::launch_image(L, n, d_bonds, d_img, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "image_kernel launch");
    return WORM_OK;
}

int worm_block(int W, int64_t n, const int32_t *d_in, int32_t *d_out, int double_bonds_value, int variant, void *stream)
{
    if (W < 2 || (W & 1) || W > 65536) return fail(WORM_E_ARG, "W = %d must be even and in [2, 65536]", W);
    if (variant != 0 && variant != 1) return fail(WORM_E_ARG, "variant = %d (0 or 1)", variant);
    if (n < 0 || (n > 0 && (!d_in || !d_out))) return fail(WORM_E_ARG, "bad n / null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    cudaError_t e = worm::launch_block(W, n, d_in, d_out, double_bonds_value, variant, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "block_kernel launch");
    return WORM_OK;
}

int worm_count(int W, int64_t n, const int32_t *d_img, int64_t *d_count, void *stream)
{
    if (W < 1 || W > 65536) return fail(WORM_E_ARG, "W = %d out of range", W);
    if (n < 0 || (n > 0 && (!d_img || !d_count))) return fail(WORM_E_ARG, "bad n / null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    cudaError_t e = worm::launch_count(W, n, d_img, d_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "count_kernel launch");
    return WORM_OK;
}

int worm_pipeline(int L, int64_t n, const uint8_t *d_bonds, int levels, int64_t *d_counts,
                  int32_t *const *h_level_images, void *stream)
{
    if (L < 4 || L > 1024 || (L & (L - 1))) return fail(WORM_E_ARG, "L = %d must be a power of two in [4, 1024]", L);
    int maxlev = 0;
    while ((L >> maxlev) >= 2) ++maxlev;
    if (levels < 1 || levels > maxlev) return fail(WORM_E_ARG, "levels = %d out of range [1, %d] for L = %d", levels, maxlev, L);
    if (n < 0 || (n > 0 && (!d_bonds || !d_counts))) return fail(WORM_E_ARG, "bad n / null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    cudaError_t e = worm::launch_pipeline(L, n, d_bonds, levels, d_counts, h_level_images, dev.smem_optin, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "pipeline_kernel launch");
    return WORM_OK;
}

int worm_boundaries(int Nx, int Ny, int64_t n, const double *d_img, int m, const double *d_cutoffs,
                    int32_t *d_links, void *stream)
{
    if (Nx < 1 || Ny < 1 || Nx > 32768 || Ny > 32768) return fail(WORM_E_ARG, "image size %d x %d out of range", Nx, Ny);
    if (n < 0 || m < 0 || ((n > 0 && m > 0) && (!d_img || !d_cutoffs || !d_links))) return fail(WORM_E_ARG, "bad n / m / null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    cudaError_t e = worm::launch_boundaries(Nx, Ny, n, m, d_img, d_cutoffs, d_links, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "boundaries_kernel launch");
    return WORM_OK;
}

static int pca_check(int d, int64_t n, int num_blocks)
{
    if (d < 1 || d > (1 << 20)) return fail(WORM_E_ARG, "d = %d out of range", d);
    if (num_blocks < 1 || num_blocks > worm::PCA_MAX_BLOCKS) return fail(WORM_E_ARG, "num_blocks = %d out of range [1, %d]", num_blocks, worm::PCA_MAX_BLOCKS);
    if (n < num_blocks || n > 0x7fffffffll) return fail(WORM_E_ARG, "n = %lld must be in [num_blocks, 2^31)", (long long)n);
    return WORM_OK;
}

size_t worm_pca_workspace_bytes(int d, int64_t n, int num_blocks, int gram_only)
{
    if (pca_check(d, n, num_blocks)) return 0;
    return worm::pca_workspace_bytes(d, n, num_blocks, !gram_only);
}

void worm_pca_timing(int enable) { worm::pca_timing(enable); }

void worm_pca_last_timing(double *ms4)
{
    double ms[worm::PCA_TIMING_SLOTS];
    worm::pca_last_timing(ms);
    if (ms4)
        for (int i = 0; i < 4; ++i) ms4[i] = ms[i];
}

int worm_pca_gram(int d, int64_t n, const int32_t *d_images, int num_blocks, int32_t *d_gram,
                  void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = pca_check(d, n, num_blocks)) return rc;
    if (!d_images || !d_gram || !d_workspace) return fail(WORM_E_ARG, "null pointer");
    if (workspace_bytes < worm::pca_workspace_bytes(d, n, num_blocks, false))
        return fail(WORM_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, worm::pca_workspace_bytes(d, n, num_blocks, false));
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const char *why = nullptr;
    cudaError_t e = worm::launch_pca_gram(d, n, d_images, num_blocks, d_gram, d_workspace, dev.sm, (cudaStream_t)stream, &why);
    if (e != cudaSuccess) return why ? fail(e == cudaErrorInvalidValue ? WORM_E_ARG : WORM_E_CUDA, "%s", why) : cuda_fail(e, "pca gram");
    return WORM_OK;
}

int worm_pca(int d, int64_t n, const int32_t *d_images, int num_blocks, int max_iters, double tol,
             double *d_explained, double *d_sigma2, double *d_component, int32_t *h_iters,
             double *h_residual, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (int rc = pca_check(d, n, num_blocks)) return rc;
    if (!d_images || !d_explained || !d_workspace) return fail(WORM_E_ARG, "null pointer");
    if (max_iters < 1 || !(tol > 0.0)) return fail(WORM_E_ARG, "max_iters >= 1 and tol > 0 required");
    if (workspace_bytes < worm::pca_workspace_bytes(d, n, num_blocks, true))
        return fail(WORM_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, worm::pca_workspace_bytes(d, n, num_blocks, true));
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const char *why = nullptr;
    int iters = 0;
    cudaError_t e = worm::launch_pca(d, n, d_images, num_blocks, max_iters, tol, d_explained, d_sigma2, d_component, &iters,
                                     h_residual, d_workspace, dev.sm, (cudaStream_t)stream, &why);
    if (e != cudaSuccess) return why ? fail(e == cudaErrorInvalidValue ? WORM_E_ARG : WORM_E_CUDA, "%s", why) : cuda_fail(e, "pca");
    if (h_iters) *h_iters = iters;
    return WORM_OK;
}

/* ------------------------------------------------------------------------------------------- */
namespace {
struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
};
}  // namespace

int worm_sim_host(int L, int64_t n_chains, const double *h_T, const double *h_seed_f64,
                  const uint64_t *h_seed_u64, uint64_t num_steps, uint64_t therm_steps,
                  int rng_mode, int algo, int64_t *h_out, uint8_t *h_bonds)
{
    if (n_chains <= 0 || !h_T || !h_out) return fail(WORM_E_ARG, "bad n_chains / null pointer");
    if (L < 2 || L > 1024) return fail(WORM_E_ARG, "L = %d out of range", L);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const size_t nb2 = 2 * (size_t)L * L;
    DevBuf bonds, out, ws, head, tail;
    cudaError_t e;
    if ((e = bonds.alloc((size_t)n_chains * nb2)) != cudaSuccess) return cuda_fail(e, "cudaMalloc bonds");
    if ((e = out.alloc((size_t)n_chains * 8 * sizeof(int64_t))) != cudaSuccess) return cuda_fail(e, "cudaMalloc out");
    int rc;
    if (rng_mode == 0) {
        if (!h_seed_f64) return fail(WORM_E_ARG, "h_seed_f64 required for rng_mode 0");
        const size_t wsb = worm_mt_workspace_bytes(n_chains);
        if ((e = ws.alloc(wsb)) != cudaSuccess) return cuda_fail(e, "cudaMalloc workspace");
        rc = worm_mt_run(L, n_chains, h_T, h_seed_f64, num_steps, 0, (uint8_t *)bonds.p, (int64_t *)out.p,
                         0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, ws.p, wsb, nullptr);
    } else if (rng_mode == 1) {
        if (!h_seed_u64) return fail(WORM_E_ARG, "h_seed_u64 required for rng_mode 1");
        const size_t wsb = worm_philox_workspace_bytes(n_chains);
        if ((e = ws.alloc(wsb)) != cudaSuccess) return cuda_fail(e, "cudaMalloc workspace");
        if ((e = head.alloc((size_t)n_chains * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc head");
        if ((e = tail.alloc((size_t)n_chains * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc tail");
        cudaMemsetAsync(bonds.p, 0, (size_t)n_chains * nb2, nullptr);
        cudaMemsetAsync(out.p, 0, (size_t)n_chains * 8 * sizeof(int64_t), nullptr);
        cudaMemsetAsync(head.p, 0, (size_t)n_chains * 4, nullptr);
        cudaMemsetAsync(tail.p, 0, (size_t)n_chains * 4, nullptr);
        rc = worm_philox_run(L, algo, n_chains, 0, h_T, h_seed_u64, 0, num_steps, therm_steps,
                             (uint8_t *)bonds.p, (int32_t *)head.p, (int32_t *)tail.p, (int64_t *)out.p,
                             nullptr, 0, nullptr, nullptr, 0, 0, nullptr, nullptr, nullptr, 0, ws.p, wsb, nullptr);
    } else {
        return fail(WORM_E_ARG, "rng_mode = %d (0 = MT19937, 1 = Philox)", rng_mode);
    }
    if (rc != WORM_OK) return rc;
    if ((e = cudaMemcpy(h_out, out.p, (size_t)n_chains * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost)) != cudaSuccess)
        return cuda_fail(e, "copy results");
    if (h_bonds && (e = cudaMemcpy(h_bonds, bonds.p, (size_t)n_chains * nb2, cudaMemcpyDeviceToHost)) != cudaSuccess)
        return cuda_fail(e, "copy bonds");
    return WORM_OK;
}

int worm_image_host(int L, int64_t n, const uint8_t *h_bonds, int32_t *h_img)
{
    if (n <= 0 || !h_bonds || !h_img) return fail(WORM_E_ARG, "bad n / null pointer");
    if (L < 3 || L > 256) return fail(WORM_E_ARG, "L = %d out of range [3, 256] for imaging", L);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const size_t nb2 = 2 * (size_t)L * L, px = 4 * (size_t)L * L;
    DevBuf in, out;
    cudaError_t e;
    if ((e = in.alloc((size_t)n * nb2)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = out.alloc((size_t)n * px * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(in.p, h_bonds, (size_t)n * nb2, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if (int rc = worm_image(L, n, (const uint8_t *)in.p, (int32_t *)out.p, nullptr)) return rc;
    if ((e = cudaMemcpy(h_img, out.p, (size_t)n * px * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    return WORM_OK;
}

int worm_block_host(int W, int64_t n, const int32_t *h_in, int32_t *h_out, int double_bonds_value, int variant)
{
    if (n <= 0 || !h_in || !h_out) return fail(WORM_E_ARG, "bad n / null pointer");
    if (W < 2 || (W & 1) || W > 65536) return fail(WORM_E_ARG, "W = %d must be even and in [2, 65536]", W);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const size_t pin = (size_t)W * W, pout = (size_t)(W / 2) * (W / 2);
    DevBuf in, out;
    cudaError_t e;
    if ((e = in.alloc((size_t)n * pin * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = out.alloc((size_t)n * pout * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(in.p, h_in, (size_t)n * pin * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if (int rc = worm_block(W, n, (const int32_t *)in.p, (int32_t *)out.p, double_bonds_value, variant, nullptr)) return rc;
    if ((e = cudaMemcpy(h_out, out.p, (size_t)n * pout * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    return WORM_OK;
}

int worm_pipeline_host(int L, int64_t n, const uint8_t *h_bonds, int levels, int64_t *h_counts)
{
    if (n < 0 || (n > 0 && (!h_bonds || !h_counts))) return fail(WORM_E_ARG, "bad n / null pointer");
    if (n == 0) return WORM_OK;
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    const size_t in_b = (size_t)n * 2 * L * L, out_b = (size_t)n * (levels > 0 ? levels : 1) * sizeof(int64_t);
    DevBuf in, out;
    cudaError_t e;
    if ((e = in.alloc(in_b)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = out.alloc(out_b)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(in.p, h_bonds, in_b, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "H2D");
    if (int rc = worm_pipeline(L, n, (const uint8_t *)in.p, levels, (int64_t *)out.p, nullptr, nullptr)) return rc;
    if ((e = cudaMemcpy(h_counts, out.p, out_b, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "D2H");
    return WORM_OK;
}

int worm_count_host(int W, int64_t n, const int32_t *h_img, int64_t *h_count)
{
    if (n <= 0 || !h_img || !h_count) return fail(WORM_E_ARG, "bad n / null pointer");
    if (W < 1 || W > 65536) return fail(WORM_E_ARG, "W = %d out of range", W);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    DevBuf in, out;
    cudaError_t e;
    if ((e = in.alloc((size_t)n * W * W * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = out.alloc((size_t)n * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(in.p, h_img, (size_t)n * W * W * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if (int rc = worm_count(W, n, (const int32_t *)in.p, (int64_t *)out.p, nullptr)) return rc;
    if ((e = cudaMemcpy(h_count, out.p, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    return WORM_OK;
}

int worm_boundaries_host(int Nx, int Ny, int64_t n, const double *h_img, int m, const double *h_cutoffs,
                         int32_t *h_links)
{
    if (n <= 0 || m <= 0 || !h_img || !h_cutoffs || !h_links) return fail(WORM_E_ARG, "bad n / m / null pointer");
    if (Nx < 1 || Ny < 1 || Nx > 32768 || Ny > 32768) return fail(WORM_E_ARG, "image size %d x %d out of range", Nx, Ny);
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    DevBuf in, cut, out;
    cudaError_t e;
    const size_t sites = (size_t)Nx * Ny;
    if ((e = in.alloc(n * sites * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cut.alloc((size_t)m * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = out.alloc((size_t)n * m * sites * 16)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = cudaMemcpy(in.p, h_img, n * sites * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if ((e = cudaMemcpy(cut.p, h_cutoffs, (size_t)m * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if (int rc = worm_boundaries(Nx, Ny, n, (const double *)in.p, m, (const double *)cut.p, (int32_t *)out.p, nullptr)) return rc;
    if ((e = cudaMemcpy(h_links, out.p, (size_t)n * m * sites * 16, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    return WORM_OK;
}

int worm_pca_host(int d, int64_t n, const int32_t *h_images, int num_blocks, int max_iters, double tol,
                  double *h_explained, double *h_component, int32_t *h_iters, double *h_residual)
{
    if (int rc = pca_check(d, n, num_blocks)) return rc;
    if (!h_images || !h_explained) return fail(WORM_E_ARG, "null pointer");
    DevInfo dev;
    if (int rc = current_device_info(dev)) return rc;
    DevBuf in, ev, comp, ws;
    cudaError_t e;
    const size_t wsb = worm::pca_workspace_bytes(d, n, num_blocks, true);
    if ((e = in.alloc((size_t)n * d * 4)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = ev.alloc((size_t)(num_blocks + 1) * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = comp.alloc((size_t)d * 8)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    if ((e = ws.alloc(wsb)) != cudaSuccess) return cuda_fail(e, "cudaMalloc workspace");
    if ((e = cudaMemcpy(in.p, h_images, (size_t)n * d * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return cuda_fail(e, "copy in");
    if (int rc = worm_pca(d, n, (const int32_t *)in.p, num_blocks, max_iters, tol, (double *)ev.p, nullptr,
                          h_component ? (double *)comp.p : nullptr, h_iters, h_residual, ws.p, wsb, nullptr))
        return rc;
    if ((e = cudaMemcpy(h_explained, ev.p, (size_t)(num_blocks + 1) * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    if (h_component && (e = cudaMemcpy(h_component, comp.p, (size_t)d * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return cuda_fail(e, "copy out");
    return WORM_OK;
}

}  // extern "C"
