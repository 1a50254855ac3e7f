// pca.cu -- leading principal component of a stack of configuration images, with the delete-one-block
// jackknife of the reference (worm_algorithm/pca.py:97-147: TruncatedSVD(n_components=1) of the
// un-centred [n_cfg, d] image matrix, explained_variance_ = var(X v1); utils.py:90-104 block_resampling
// = KFold(num_blocks) training sets; utils.py:123-135 jackknife_err).
//
// Device pipeline (everything stays in HBM):
//   1. pca_pack_kernel   int32 images [n, d] -> int8 Xt [d, Ktot], pixel-major, each jackknife block in its
//                        own zero-padded run of 128-wide k tiles; per-block column sums; max |value|.
//   2. gram_kernel       per-block Gram matrices G_b = X_b^T X_b on the 5th-generation tensor cores
//                        (tcgen05.mma kind::i8, s8 x s8 -> s32 in TMEM; TMA 128B-swizzled operand
//                        tiles; only the lower triangle of 256x256 tiles is computed and mirrored on the
//                        way out).  Integer in, integer out: the result is EXACT as long as it fits
//                        int32 (checked on the host from max|value| and n; larger inputs are refused).
//                        worm_pca needs only the Gram matrix of ALL configurations (one "block" spanning
//                        the whole packed matrix); worm_pca_gram exposes the per-block matrices.
//   3. power iteration   for the 1 + num_blocks problems at once (full data and each delete-one
//                        set): A_j v_j = G_full v_j - X_b^T (X_b v_j), b = j - 1.  G_full meets all
//                        vectors in one pass (pca_matvec_sym_kernel: G is symmetric, so a lane owning row r walks the
//                        coalesced column r of the rows above it); the deleted block's part goes
//                        through the int8 configurations themselves (pca_xv_kernel, pca_xty_kernel:
//                        two passes over 1 byte per pixel instead of a d x d matrix per block).
//                        Matrices are exact integers, vectors and accumulation fp64.
//   4. explained variance lambda_j / n_j - (s_j . v_j / n_j)^2.
#include <cuda.h>   // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through the runtime
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <vector>

#include "worm_kernels.h"

namespace worm {

namespace {

// ---------------------------------------------------------------------------------------------------
// 1. pack: transpose + convert, block column sums
// ---------------------------------------------------------------------------------------------------
struct PackPlan {
    int d, nb;
    int64_t ktot;
    int cfg0[PCA_MAX_BLOCKS + 1];   // first configuration of jackknife block b (cfg0[nb] = n)
    int kt0[PCA_MAX_BLOCKS + 1];    // first 64-configuration group of block b in Xt (kt0[nb] = Ktot / 64); multiples of 8
};

__global__ void __launch_bounds__(256) pca_pack_kernel(const int32_t *__restrict__ img, int8_t *__restrict__ xt,
                                                       long long *__restrict__ colsum, int *__restrict__ maxabs,
                                                       const PackPlan p)
{
    __shared__ int tile[64][65];
    __shared__ int csum[64];
    const int kt = blockIdx.x, c0 = blockIdx.y * 64, t = threadIdx.x;
    int b = 0;
    while (kt >= p.kt0[b + 1]) ++b;
    const int cfg_first = p.cfg0[b] + (kt - p.kt0[b]) * 64, cfg_end = p.cfg0[b + 1];
    if (t < 64) csum[t] = 0;
    int mx = 0, part = 0;
    {
        const int c = c0 + (t & 63);
        int v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {   // sixteen independent loads in flight per thread
            const int cfg = cfg_first + (t >> 6) + 4 * i;
            v[i] = (cfg < cfg_end && c < p.d) ? __ldcs(img + (size_t)cfg * p.d + c) : 0;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            tile[(t >> 6) + 4 * i][t & 63] = v[i];
            mx = max(mx, abs(v[i]));
            part += v[i];
        }
    }
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // one address for the whole grid: only touch it when this warp would raise it
    if ((t & 31) == 0 && mx > *reinterpret_cast<volatile int *>(maxabs)) atomicMax(maxabs, mx);
    __syncthreads();
    if (part) atomicAdd(&csum[t & 63], part);
    {
        const int k4 = t & 15;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int cl = (t >> 4) + 16 * i, c = c0 + cl;
            if (c < p.d) {
                const uint32_t w = (uint32_t)(tile[4 * k4][cl] & 0xff) | ((uint32_t)(tile[4 * k4 + 1][cl] & 0xff) << 8) |
                                   ((uint32_t)(tile[4 * k4 + 2][cl] & 0xff) << 16) | ((uint32_t)(tile[4 * k4 + 3][cl] & 0xff) << 24);
                *reinterpret_cast<uint32_t *>(xt + (size_t)c * p.ktot + (size_t)kt * 64 + 4 * k4) = w;
            }
        }
    }
    __syncthreads();
    if (t < 64 && c0 + t < p.d && csum[t])
        atomicAdd(reinterpret_cast<unsigned long long *>(colsum + (size_t)b * p.d + c0 + t), (unsigned long long)(long long)csum[t]);
}

// ---------------------------------------------------------------------------------------------------
// 2. Gram matrices on tcgen05
// ---------------------------------------------------------------------------------------------------
// CTA tile 256 x 256: two M = 128 accumulators (TMEM columns 0-255 and 256-511) share every B stage, so a
// stage of 64 KB feeds eight MMAs (1024 tensor cycles).  With three stages in flight that covers ~3000
// cycles of TMA latency at 64 B/clk/SM of L2 traffic; the 128 x 256 tile of the first version (48 KB per
// 512 cycles, four stages) left the tensor pipe idle 47 % of the time waiting for operands.
constexpr int BM = 256, BN = 256, BK = 128, STAGES = 3;   // BK int8 elements = one 128-byte swizzle row
constexpr int HALF_BYTES = 128 * BK;                        // one TMA box: 128 rows x 128 bytes
constexpr int STAGE_BYTES = 4 * HALF_BYTES;                 // A rows 0-127, A rows 128-255, B cols 0-127, B cols 128-255
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int TMEM_COLS = 512;

struct GramPlan {
    int d, ldg, nb, tm_count, tiles_per_block;
    int kt0[PCA_MAX_BLOCKS + 1];   // first BK-wide k block of jackknife block b
    int32_t *G;   // [nb][ldg][ldg]
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A protocol error must end the launch with an error, never hang the device: the spin is bounded.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try(bar, parity))
        if (++spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte rows, 128B swizzle (what the TMA box {64 x rows} with SWIZZLE_128B
// lays down): 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr)
{
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::i8 instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t idesc_i8(int n)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Tile {
    int b, tm, tn, mh, nh;   // mh / nh: 128-row halves of the tile that hold any row / column below ldg
};
// Lower-triangle enumeration of 256 x 256 tiles: row tile tm meets column tiles tn <= tm.
__device__ __forceinline__ Tile decode_tile(int t, const GramPlan &p)
{
    Tile r;
    r.b = t / p.tiles_per_block;
    int rem = t - r.b * p.tiles_per_block, tm = 0;
    for (; rem > tm; ++tm) rem -= tm + 1;
    r.tm = tm;
    r.tn = rem;
    r.mh = p.ldg - tm * BM > 128 ? 2 : 1;
    r.nh = p.ldg - rem * BN > 128 ? 2 : 1;
    return r;
}

__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_kernel(const __grid_constant__ CUtensorMap tmap, const GramPlan p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t tfull = bars + 8u * (2 * STAGES), tempty = bars + 8u * (2 * STAGES + 1);
    const uint32_t slot = bars + 8u * (2 * STAGES + 2);
    volatile uint32_t *slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.nb * p.tiles_per_block;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(empty(s), 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;

    if (warp == 0) {
        // ===== TMA producer (runs ahead into the next tile while the epilogue drains TMEM) =====
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const Tile tl = decode_tile(t, p);
            const int k0 = p.kt0[tl.b], nk = p.kt0[tl.b + 1] - k0;
            for (int kb = 0; kb < nk; ++kb) {
                if (lane == 0) {
                    mbar_wait(empty(stage), phase ^ 1);
                    const uint32_t a = base + stage * STAGE_BYTES;
                    mbar_expect_tx(full(stage), (tl.mh + tl.nh) * HALF_BYTES);
                    const int kc = (k0 + kb) * BK;
                    tma_load_2d(a, &tmap, full(stage), kc, tl.tm * BM);
                    if (tl.mh > 1) tma_load_2d(a + HALF_BYTES, &tmap, full(stage), kc, tl.tm * BM + 128);
                    tma_load_2d(a + 2 * HALF_BYTES, &tmap, full(stage), kc, tl.tn * BN);
                    if (tl.nh > 1) tma_load_2d(a + 3 * HALF_BYTES, &tmap, full(stage), kc, tl.tn * BN + 128);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int stage = 0, phase = 0, it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const Tile tl = decode_tile(t, p);
            const int nk = p.kt0[tl.b + 1] - p.kt0[tl.b];
            const uint32_t idesc = idesc_i8(128 * tl.nh);
            if (lane == 0) {
                mbar_wait(tempty, (it & 1) ^ 1);     // the epilogue of the previous tile has drained both accumulators
                tc_fence_after();
            }
            __syncwarp();
            for (int kb = 0; kb < nk; ++kb) {
                if (lane == 0) {
                    mbar_wait(full(stage), phase);
                    tc_fence_after();
                    const uint32_t a = base + stage * STAGE_BYTES;
                    const uint64_t a0 = smem_desc_sw128(a), a1 = smem_desc_sw128(a + HALF_BYTES);
                    const uint64_t bd = smem_desc_sw128(a + 2 * HALF_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 32; ++k) {   // one instruction consumes K = 32 bytes
                        umma_i8(tmem, a0 + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                        if (tl.mh > 1) umma_i8(tmem + 256, a1 + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(empty(stage));
                    if (kb == nk - 1) umma_commit(tfull);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> G (direct rows, and the mirrored transpose below the diagonal) =====
        const int w = warp - 4;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const Tile tl = decode_tile(t, p);
            mbar_wait(tfull, it & 1);
            tc_fence_after();
            int32_t *Gb = p.G + (size_t)tl.b * p.ldg * p.ldg;
            const bool mirror = tl.tn < tl.tm;
            for (int h = 0; h < tl.mh; ++h) {
                const int row = tl.tm * BM + 128 * h + 32 * w + lane;
                for (int c0 = 0; c0 < 128 * tl.nh; c0 += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem + (uint32_t)(256 * h + c0) + ((uint32_t)(32 * w) << 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
                        "tcgen05.wait::ld.sync.aligned;"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr)
                        : "memory");
                    const int col0 = tl.tn * BN + c0;
                    if (row < p.ldg) {
                        int32_t *dst = Gb + (size_t)row * p.ldg + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (col0 + j < p.ldg) *reinterpret_cast<uint4 *>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (mirror) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.ldg) Gb[(size_t)(col0 + j) * p.ldg + row] = (int32_t)v[j];
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// 3. power iteration on the nj = 1 + nb problems A_0 = G_full, A_j = G_full - X_{j-1}^T X_{j-1}
// ---------------------------------------------------------------------------------------------------
constexpr int MV_BATCH = 4;   // iterations queued per host round trip

// iteration state shared by the kernels of one worm_pca call
struct IterState {
    int done;             // set once every problem has converged; later launches return at once
    int iters;            // iterations actually executed
    int cnt[MV_BATCH];    // cnt[i]: problems whose |dv|^2 <= tol^2 at iteration i of the current batch
};

__device__ __forceinline__ double warp_sum(double x)
{
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

__device__ __forceinline__ bool iter_finished(const IterState *s, int slot, int nj)
{
    return s->done || (slot > 0 && s->cnt[slot - 1] == nj);
}

// int32 -> double through the mantissa of 2^52 (exact for |g| < 2^31): one fp64 add instead of a
// conversion instruction.
__device__ __forceinline__ double int_to_double(int g)
{
    return __hiloint2double(0x43300000, g ^ 0x80000000) - 4503601774854144.0;   // 2^52 + 2^31
}

// Signed byte B of a packed word -> double, exactly, in three instructions per element (permute, fp64 add,
// then the FMA that uses it): with wx = word ^ 0x80808080 every byte is u = x + 128 in [0, 255]; the permute
// drops u into the low word of 2^52 (= 2^52 + u exactly) and the add removes 2^52 + 128.
template <int B>
__device__ __forceinline__ double s8_to_double(uint32_t wx)
{
    return __hiloint2double(0x43300000, (int)__byte_perm(wx, 0u, 0x4440u | B)) - 4503599627370624.0;
}

// first 512-configuration step of each jackknife block in the packed matrix
struct StepPlan {
    int nb;
    int st0[PCA_MAX_BLOCKS + 1];
};

constexpr int XV_SLICES = 16;  // pixel-row slices per 512-configuration step (grid.y of pca_xv_kernel)

// y = X_b v_{b+1} for every block b at once: y[k] = sum_c Xt[c][k] * V_{b(k)+1}[c].
// blockIdx.x = 512-configuration step (one jackknife block per step), blockIdx.y = slice of pixel rows;
// a warp reads 512 contiguous bytes of a pixel row, lane l owning configurations 16 l .. 16 l + 15.
// The result is accumulated (fp64 atomics across the XV_SLICES slices) in the lane-major order
// ysw[((step * 8 + i) * 32 + l) * 2 + e] = y[step * 512 + 16 l + 2 i + e], which pca_xty_kernel reads back
// with coalesced 16-byte loads.
__global__ void __launch_bounds__(256) pca_xv_kernel(const int8_t *__restrict__ xt, const double *__restrict__ V,
                                                     double *__restrict__ ysw, int d, int ldg, long long ktot,
                                                     const StepPlan sp, const IterState *state, int slot)
{
    __shared__ double red[8][16][32];
    if (iter_finished(state, slot, sp.nb + 1)) return;
    const int step = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int b = 0;
    while (step >= sp.st0[b + 1]) ++b;
    const double *v = V + (size_t)(b + 1) * ldg;
    const int rps = (d + XV_SLICES - 1) / XV_SLICES;
    const int c_begin = blockIdx.y * rps, c_end = min(d, c_begin + rps);
    const int8_t *col = xt + (size_t)step * 512 + lane * 16;
    double acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
    constexpr int U = 4;
    // software pipeline: the loads of the next batch of rows are in flight while this one is consumed
    auto load = [&](int c, uint4 (&x)[U], double (&vc)[U]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int cc = c + 8 * u;
            const bool ok = cc < c_end;
            x[u] = ok ? *reinterpret_cast<const uint4 *>(col + (size_t)cc * ktot) : make_uint4(0u, 0u, 0u, 0u);
            vc[u] = ok ? v[cc] : 0.0;
        }
    };
    uint4 x[U], xn[U];
    double vc[U], vn[U];
    load(c_begin + warp, x, vc);
    for (int c = c_begin + warp; c < c_end; c += 8 * U) {
        load(c + 8 * U, xn, vn);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t w[4] = {x[u].x ^ 0x80808080u, x[u].y ^ 0x80808080u, x[u].z ^ 0x80808080u, x[u].w ^ 0x80808080u};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[4 * q + 0] = fma(s8_to_double<0>(w[q]), vc[u], acc[4 * q + 0]);
                acc[4 * q + 1] = fma(s8_to_double<1>(w[q]), vc[u], acc[4 * q + 1]);
                acc[4 * q + 2] = fma(s8_to_double<2>(w[q]), vc[u], acc[4 * q + 2]);
                acc[4 * q + 3] = fma(s8_to_double<3>(w[q]), vc[u], acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            x[u] = xn[u];
            vc[u] = vn[u];
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) red[warp][e][lane] = acc[e];
    __syncthreads();
    {
        const int l = threadIdx.x >> 3, i = threadIdx.x & 7;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            s0 += red[w][2 * i][l];
            s1 += red[w][2 * i + 1][l];
        }
        double *dst = ysw + ((size_t)(step * 8 + i) * 32 + l) * 2;
        atomicAdd(dst, s0);
        atomicAdd(dst + 1, s1);
    }
}

// WB_{b+1} = X_b^T y_b: WB[(b+1)][c] = sum_{k in block b} Xt[c][k] * y[k].  blockIdx.x = 32 pixel rows
// (4 per warp), blockIdx.y = jackknife block.
__global__ void __launch_bounds__(256) pca_xty_kernel(const int8_t *__restrict__ xt, const double *__restrict__ ysw,
                                                      double *__restrict__ WB, int d, int ldg, long long ktot,
                                                      const StepPlan sp, const IterState *state, int slot)
{
    if (iter_finished(state, slot, sp.nb + 1)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, c0 = blockIdx.x * 32 + warp * 4;
    if (c0 >= ldg) return;
    const double2 *y2 = reinterpret_cast<const double2 *>(ysw);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    auto load_x = [&](int step, uint4 (&x)[4]) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            x[r] = (c0 + r < d && step < sp.st0[b + 1])
                       ? *reinterpret_cast<const uint4 *>(xt + (size_t)(c0 + r) * ktot + (size_t)step * 512 + lane * 16)
                       : make_uint4(0u, 0u, 0u, 0u);
    };
    uint4 x[4], xn[4];
    load_x(sp.st0[b], x);
    for (int step = sp.st0[b]; step < sp.st0[b + 1]; ++step) {
        double2 y[8];
        load_x(step + 1, xn);                                  // in flight while this step is consumed
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = y2[(size_t)(step * 8 + i) * 32 + lane];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t w[4] = {x[r].x ^ 0x80808080u, x[r].y ^ 0x80808080u, x[r].z ^ 0x80808080u, x[r].w ^ 0x80808080u};
            double a = acc[r];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a = fma(s8_to_double<0>(w[q]), y[2 * q].x, a);
                a = fma(s8_to_double<1>(w[q]), y[2 * q].y, a);
                a = fma(s8_to_double<2>(w[q]), y[2 * q + 1].x, a);
                a = fma(s8_to_double<3>(w[q]), y[2 * q + 1].y, a);
            }
            acc[r] = a;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) x[r] = xn[r];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double sum = warp_sum(acc[r]);
        if (lane == 0 && c0 + r < ldg) WB[(size_t)(b + 1) * ldg + c0 + r] = c0 + r < d ? sum : 0.0;
    }
}

// WA_j += G V_j for up to SYM_J problems, using the symmetry of G: lane l of a warp owns rows r0 + 2l and r0 + 2l + 1
// and walks down COLUMNS of those rows, i.e. along row c of G (G[c][r] = G[r][c]) -- a coalesced 128-byte
// load per c, while V_j[c] is the same for every lane (one broadcast load of the transposed copy VT[c][j]).
// No shared-memory staging of the vectors, no cross-lane reduction; the eight warps of a block take
// interleaved c and are summed through shared memory, the SYM_SPLIT column ranges through fp64 atomics
// (WA is zeroed before the launch).  fp64 issue (32 lanes/clk/SM) is the bound: nj d^2 FMAs.
constexpr int SYM_J = 12, SYM_SPLIT = 4;

__global__ void __launch_bounds__(256, 2) pca_matvec_sym_kernel(const int32_t *__restrict__ G, const double *__restrict__ VT,
                                                                 double *__restrict__ WA, int ldg, int nj,
                                                                 const IterState *state, int slot)
{
    __shared__ double red[8][2 * SYM_J][32];
    if (iter_finished(state, slot, nj)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 64, ra = r0 + 2 * lane;          // this lane's rows: ra, ra + 1 (one 8-byte load per matrix row)
    const int cw = ((ldg + SYM_SPLIT - 1) / SYM_SPLIT + 7) & ~7;
    const int cb = blockIdx.y * cw, ce = min(ldg, cb + cw);
    const int j0 = blockIdx.z * SYM_J, jn = min(SYM_J, nj - j0);
    const double2 *vt = reinterpret_cast<const double2 *>(VT + (size_t)blockIdx.z * ldg * SYM_J);
    const bool oka = ra < ldg;                                    // ldg is even: ra + 1 < ldg too
    double acc[2][SYM_J];
#pragma unroll
    for (int j = 0; j < SYM_J; ++j) acc[0][j] = acc[1][j] = 0.0;
    constexpr int U = 8;
    for (int c = cb + warp; c < ce; c += 8 * U) {
        int2 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int cc = c + 8 * u;
            g[u] = (cc < ce && oka) ? __ldcs(reinterpret_cast<const int2 *>(G + (size_t)cc * ldg + ra)) : make_int2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int cc = min(c + 8 * u, ce - 1);
            const double da = int_to_double(g[u].x), db = int_to_double(g[u].y);
#pragma unroll
            for (int jj = 0; jj < SYM_J / 2; ++jj) {
                const double2 v = vt[(size_t)cc * (SYM_J / 2) + jj];
                acc[0][2 * jj] = fma(da, v.x, acc[0][2 * jj]);
                acc[0][2 * jj + 1] = fma(da, v.y, acc[0][2 * jj + 1]);
                acc[1][2 * jj] = fma(db, v.x, acc[1][2 * jj]);
                acc[1][2 * jj + 1] = fma(db, v.y, acc[1][2 * jj + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < SYM_J; ++j) {
        red[warp][j][lane] = acc[0][j];
        red[warp][SYM_J + j][lane] = acc[1][j];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * SYM_J * 32; o += 256) {
        const int l = o & 31, k = o >> 5;                       // k = half * SYM_J + j
        const int j = k % SYM_J, row = r0 + 2 * l + k / SYM_J;
        if (j < jn && row < ldg) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += red[w][k][l];
            atomicAdd(WA + (size_t)(j0 + j) * ldg + row, sum);
        }
    }
}

__device__ double block_sum(double x, double *scratch)
{
    x = warp_sum(x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += scratch[i];
    return s;
}

// start vectors: the normalised column sums of the included configurations (for images, close to the
// leading singular vector already) plus a small fixed pattern so that no start is exactly zero.
__global__ void __launch_bounds__(256) pca_coltot_kernel(const long long *__restrict__ colsum, long long *__restrict__ coltot, int d, int nb)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    long long s = 0;
    for (int b = 0; b < nb; ++b) s += colsum[(size_t)b * d + c];
    coltot[c] = s;
}

__global__ void __launch_bounds__(1024) pca_init_kernel(const long long *__restrict__ colsum, const long long *__restrict__ coltot,
                                                       double *__restrict__ V, double *__restrict__ VT, int d, int ldg)
{
    __shared__ double scratch[32];
    const int j = blockIdx.x;
    double *v = V + (size_t)j * ldg;
    double ss = 0.0;
    for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
        double x = 0.0;
        if (c < d) x = (double)(coltot[c] - (j ? colsum[(size_t)(j - 1) * d + c] : 0ll));
        v[c] = x;
        ss += x * x;
    }
    const double nrm = sqrt(block_sum(ss, scratch));
    double s2 = 0.0;
    for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
        double x = 0.0;
        if (c < d) {
            const uint32_t h = (uint32_t)c * 2654435761u + 0x9E3779B9u;
            x = (nrm > 0.0 ? v[c] / nrm : 0.0) + 1e-3 * (0.5 + (double)(h >> 8) * (1.0 / 16777216.0)) / sqrt((double)d);
        }
        v[c] = x;
        s2 += x * x;
    }
    const double n2 = sqrt(block_sum(s2, scratch));
    for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
        const double x = v[c] / n2;
        v[c] = x;
        VT[((size_t)(j / SYM_J) * ldg + c) * SYM_J + j % SYM_J] = x;
    }
}

// W_j = WA_j - WB_j;  stats[j] = { Rayleigh quotient V.W (|V| = 1), |V_new - V|^2 };  V <- W / |W|.
__global__ void __launch_bounds__(1024) pca_update_kernel(double *__restrict__ V, double *__restrict__ VT, const double *__restrict__ WA,
                                                         const double *__restrict__ WB, double *__restrict__ stats, int ldg,
                                                         int nj, double tol2, IterState *state, int slot)
{
    __shared__ double scratch[32];
    if (iter_finished(state, slot, nj)) {
        if (threadIdx.x == 0) state->done = 1;
        return;
    }
    const int j = blockIdx.x;
    double *v = V + (size_t)j * ldg;
    const double *wa = WA + (size_t)j * ldg, *wb = WB + (size_t)j * ldg;
    double vw = 0.0, ww = 0.0;
    for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
        const double w = j ? wa[c] - wb[c] : wa[c];
        vw += v[c] * w;
        ww += w * w;
    }
    vw = block_sum(vw, scratch);
    ww = block_sum(ww, scratch);
    double diff = 0.0;
    if (ww > 0.0) {
        const double inv = 1.0 / sqrt(ww);
        for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
            const double x = (j ? wa[c] - wb[c] : wa[c]) * inv, dx = x - v[c];
            diff += dx * dx;
            v[c] = x;
            VT[((size_t)(j / SYM_J) * ldg + c) * SYM_J + j % SYM_J] = x;
        }
    }
    diff = block_sum(diff, scratch);
    if (threadIdx.x == 0) {
        stats[2 * j] = vw;
        stats[2 * j + 1] = diff;
        if (diff <= tol2) atomicAdd(&state->cnt[slot], 1);
        if (j == 0) state->iters += 1;
    }
}

// explained variance of the projection on v_j: mean((X v)^2) - mean(X v)^2 over the included configurations
__global__ void __launch_bounds__(1024) pca_finish_kernel(const double *__restrict__ V, const double *__restrict__ stats,
                                                         const long long *__restrict__ colsum, const long long *__restrict__ coltot,
                                                         const PackPlan p, int ldg,
                                                         double *__restrict__ explained, double *__restrict__ sigma2,
                                                         double *__restrict__ component)
{
    __shared__ double scratch[32];
    const int j = blockIdx.x;
    const double *v = V + (size_t)j * ldg;
    double sv = 0.0, vs = 0.0;
    for (int c = threadIdx.x; c < p.d; c += blockDim.x) {
        const long long s = coltot[c] - (j ? colsum[(size_t)(j - 1) * p.d + c] : 0ll);
        sv += (double)s * v[c];
        vs += v[c];
    }
    sv = block_sum(sv, scratch);
    vs = block_sum(vs, scratch);
    const double n = (double)(p.cfg0[p.nb] - (j > 0 ? p.cfg0[j] - p.cfg0[j - 1] : 0));
    if (threadIdx.x == 0) {
        const double m = n > 0.0 ? sv / n : 0.0;
        explained[j] = n > 0.0 ? stats[2 * j] / n - m * m : 0.0;   // num_blocks = 1: the delete-one set is empty
        if (sigma2) sigma2[j] = stats[2 * j];
    }
    if (j == 0 && component) {   // sign convention: the component sums to a positive number
        const double sg = vs < 0.0 ? -1.0 : 1.0;
        for (int c = threadIdx.x; c < p.d; c += blockDim.x) component[c] = sg * v[c];
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// A second stream + fork / join events, created once per host thread and device (stream creation costs
// tens of microseconds, a worm_pca call one millisecond) and kept for the life of the thread.
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaError_t init()
    {
        if (s) return cudaSuccess;
        cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
        return e;
    }
};

SideStream *side_stream()
{
    constexpr int MAX_DEV = 64;
    thread_local SideStream cache[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    return cache[dev].init() == cudaSuccess ? &cache[dev] : nullptr;
}

// Optional stage timing (worm_pca_timing): CUDA events on the caller's stream around each stage.
std::atomic<int> g_timing{0};
thread_local double g_ms[PCA_TIMING_SLOTS] = {0, 0, 0, 0, 0};

struct StageTimer {
    cudaStream_t st;
    bool on;
    cudaEvent_t ev[PCA_TIMING_SLOTS + 1];
    int n = 0;
    explicit StageTimer(cudaStream_t s) : st(s), on(g_timing.load() != 0)
    {
        if (on)
            for (auto &e : ev) cudaEventCreate(&e);
    }
    void mark()
    {
        if (on && n <= PCA_TIMING_SLOTS) cudaEventRecord(ev[n++], st);
    }
    void finish()   // after a stream synchronise
    {
        if (!on) return;
        for (int i = 0; i < PCA_TIMING_SLOTS; ++i) {
            float ms = 0.f;
            if (i + 1 < n) cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            g_ms[i] = ms;
        }
    }
    ~StageTimer()
    {
        if (on)
            for (auto &e : ev) cudaEventDestroy(e);
    }
};

void make_pack_plan(int d, int64_t n, int nb, PackPlan &p)
{
    p.d = d;
    p.nb = nb;
    // sklearn KFold(n_splits = nb) without shuffling: the first n % nb folds hold n / nb + 1 samples
    int cfg = 0, kt = 0;
    for (int b = 0; b < nb; ++b) {
        const int len = (int)(n / nb) + (b < n % nb ? 1 : 0);
        p.cfg0[b] = cfg;
        p.kt0[b] = kt;
        cfg += len;
        kt += 8 * ((len + 511) / 512);   // 64-configuration groups; every block is padded to whole 512-configuration steps
    }
    p.cfg0[nb] = cfg;
    p.kt0[nb] = kt;
    p.ktot = (int64_t)kt * 64;
}

struct Layout {
    size_t xt, colsum, maxabs, gram, v, vt, w, wb, ysw, stats, state, coltot, total;
    int ldg;
};

Layout make_layout(int d, const PackPlan &p, bool with_iteration)
{
    Layout l;
    l.ldg = (d + 3) & ~3;
    size_t off = 0;
    l.xt = off;      off = align_up(off + (size_t)d * p.ktot, 1024);
    l.colsum = off;  off = align_up(off + (size_t)p.nb * d * 8, 256);
    l.maxabs = off;  off = align_up(off + 256, 256);
    l.gram = off;
    if (with_iteration) {
        off = align_up(off + (size_t)l.ldg * l.ldg * 4, 256);   // the Gram matrix of all configurations
        l.v = off;      off = align_up(off + (size_t)(p.nb + 1) * l.ldg * 8, 256);
        l.vt = off;     off = align_up(off + (size_t)((p.nb + SYM_J) / SYM_J) * l.ldg * SYM_J * 8, 256);
        l.w = off;      off = align_up(off + (size_t)(p.nb + 1) * l.ldg * 8, 256);
        l.wb = off;     off = align_up(off + (size_t)(p.nb + 1) * l.ldg * 8, 256);
        l.ysw = off;    off = align_up(off + (size_t)p.ktot * 8, 256);
        l.stats = off;  off = align_up(off + (size_t)(p.nb + 1) * 16, 256);
        l.state = off;  off = align_up(off + sizeof(IterState), 256);
        l.coltot = off; off = align_up(off + (size_t)d * 8, 256);
    } else {
        l.v = l.vt = l.w = l.wb = l.ysw = l.stats = l.state = l.coltot = off;
    }
    l.total = off;
    return l;
}

cudaError_t run_pack_and_gram(int d, int64_t n, const int32_t *d_img, const PackPlan &pp, const Layout &lay, uint8_t *ws,
                              int32_t *d_gram, bool all_in_one, int sm_count, cudaStream_t st, const char **why, StageTimer &tm)
{
    cudaError_t e;
    tm.mark();
    auto *xt = reinterpret_cast<int8_t *>(ws + lay.xt);
    auto *colsum = reinterpret_cast<long long *>(ws + lay.colsum);
    int *maxabs = reinterpret_cast<int *>(ws + lay.maxabs);
    if ((e = cudaMemsetAsync(ws + lay.colsum, 0, (lay.maxabs + 256) - lay.colsum, st)) != cudaSuccess) return e;
    {
        dim3 grid((unsigned)pp.kt0[pp.nb], (unsigned)((d + 63) / 64));
        pca_pack_kernel<<<grid, 256, 0, st>>>(d_img, xt, colsum, maxabs, pp);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    tm.mark();
    // exactness gate: every value must fit int8 and every entry of the full Gram matrix int32
    int h_max = 0;
    if ((e = cudaMemcpyAsync(&h_max, maxabs, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (h_max > 127 || (int64_t)h_max * h_max * n >= (1ll << 31)) {
        *why = "pixel values too large for an exact s8 x s8 -> s32 Gram matrix (need max|v| <= 127 and max|v|^2 * n < 2^31)";
        return cudaErrorInvalidValue;
    }

    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        *why = "cuTensorMapEncodeTiled not available from this driver";
        return cudaErrorNotSupported;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)pp.ktot, (cuuint64_t)d};
    const cuuint64_t gstride[1] = {(cuuint64_t)pp.ktot};
    const cuuint32_t box[2] = {BK, 128};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, xt, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        *why = "cuTensorMapEncodeTiled failed";
        return cudaErrorInvalidValue;
    }
    GramPlan gp;
    gp.d = d;
    gp.ldg = lay.ldg;
    gp.nb = all_in_one ? 1 : pp.nb;   // all_in_one: one Gram matrix over the whole packed matrix (the padding is zeros)
    gp.tm_count = (lay.ldg + BM - 1) / BM;
    gp.tiles_per_block = gp.tm_count * (gp.tm_count + 1) / 2;
    for (int b = 0; b <= pp.nb; ++b) gp.kt0[b] = pp.kt0[b] / 2;
    if (all_in_one) gp.kt0[1] = pp.kt0[pp.nb] / 2;
    gp.G = d_gram;
    // per device (the attribute lives in the device's context), so set on every call: it costs microseconds
    if ((e = cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM)) != cudaSuccess) return e;
    const int total = gp.nb * gp.tiles_per_block;
    tm.mark();
    gram_kernel<<<std::min(total, sm_count), GRAM_THREADS, GRAM_SMEM, st>>>(tmap, gp);
    tm.mark();
    return cudaGetLastError();
}

}  // namespace

size_t pca_workspace_bytes(int d, int64_t n, int nb, bool with_iteration)
{
    PackPlan pp;
    make_pack_plan(d, n, nb, pp);
    return make_layout(d, pp, with_iteration).total;
}

int pca_ldg(int d) { return (d + 3) & ~3; }

cudaError_t launch_pca_gram(int d, int64_t n, const int32_t *d_img, int nb, int32_t *d_gram, void *d_ws, int sm_count,
                            cudaStream_t st, const char **why)
{
    PackPlan pp;
    make_pack_plan(d, n, nb, pp);
    const Layout lay = make_layout(d, pp, false);
    StageTimer tm(st);
    cudaError_t e = run_pack_and_gram(d, n, d_img, pp, lay, static_cast<uint8_t *>(d_ws), d_gram, false, sm_count, st, why, tm);
    if (e == cudaSuccess && tm.on) {
        e = cudaStreamSynchronize(st);
        tm.finish();
    }
    return e;
}

void pca_timing(int enable) { g_timing.store(enable); }

void pca_last_timing(double *ms)
{
    for (int i = 0; i < PCA_TIMING_SLOTS; ++i) ms[i] = g_ms[i];
}

cudaError_t launch_pca(int d, int64_t n, const int32_t *d_img, int nb, int max_iters, double tol, double *d_explained,
                       double *d_sigma2, double *d_component, int *h_iters, double *h_residual, void *d_ws, int sm_count,
                       cudaStream_t st, const char **why)
{
    PackPlan pp;
    make_pack_plan(d, n, nb, pp);
    const Layout lay = make_layout(d, pp, true);
    uint8_t *ws = static_cast<uint8_t *>(d_ws);
    int32_t *G = reinterpret_cast<int32_t *>(ws + lay.gram);
    StageTimer tm(st);
    cudaError_t e = run_pack_and_gram(d, n, d_img, pp, lay, ws, G, true, sm_count, st, why, tm);
    if (e != cudaSuccess) return e;
    const int ldg = lay.ldg, nj = nb + 1;
    auto *colsum = reinterpret_cast<const long long *>(ws + lay.colsum);
    double *V = reinterpret_cast<double *>(ws + lay.v), *W = reinterpret_cast<double *>(ws + lay.w);
    double *stats = reinterpret_cast<double *>(ws + lay.stats);
    long long *coltot = reinterpret_cast<long long *>(ws + lay.coltot);
    pca_coltot_kernel<<<(d + 255) / 256, 256, 0, st>>>(colsum, coltot, d, nb);
    const int vt = ldg >= 2048 ? 1024 : 256;   // threads of the per-problem vector kernels (latency bound: one block per problem)
    double *VT = reinterpret_cast<double *>(ws + lay.vt);
    const int njc = (nj + SYM_J - 1) / SYM_J;
    if ((e = cudaMemsetAsync(VT, 0, (size_t)njc * ldg * SYM_J * 8, st)) != cudaSuccess) return e;   // unused problem slots stay zero
    pca_init_kernel<<<nj, vt, 0, st>>>(colsum, coltot, V, VT, d, ldg);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;

    double *WB = reinterpret_cast<double *>(ws + lay.wb);
    IterState *state = reinterpret_cast<IterState *>(ws + lay.state);
    if ((e = cudaMemsetAsync(state, 0, sizeof(IterState), st)) != cudaSuccess) return e;
    const int8_t *xt = reinterpret_cast<const int8_t *>(ws + lay.xt);
    double *ysw = reinterpret_cast<double *>(ws + lay.ysw);
    StepPlan sp;
    sp.nb = nb;
    for (int b = 0; b <= nb; ++b) sp.st0[b] = pp.kt0[b] / 8;
    std::vector<double> h_stats(2 * (size_t)nj);
    IterState h_state{};
    int iters = 0;
    double worst = 0.0;
    // Batches of MV_BATCH iterations are queued without host round trips; once every problem has
    // converged (device-side count) the remaining launches of the batch return immediately.
    // The two halves of an iteration are independent (G v on one side, the deleted blocks' X^T X v on the
    // other): the second runs on a side stream, forked and joined with events around every update.
    SideStream *sidep = side_stream();
    if (!sidep) {
        *why = "could not create the side stream";
        return cudaErrorUnknown;
    }
    SideStream &side = *sidep;
    while (iters < max_iters) {
        const int batch = std::min(MV_BATCH, max_iters - iters);
        if ((e = cudaMemsetAsync(state->cnt, 0, sizeof(state->cnt), st)) != cudaSuccess) return e;
        for (int i = 0; i < batch; ++i) {
            if ((e = cudaEventRecord(side.fork, st)) != cudaSuccess) return e;
            if ((e = cudaStreamWaitEvent(side.s, side.fork, 0)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(ysw, 0, (size_t)pp.ktot * 8, side.s)) != cudaSuccess) return e;
            pca_xv_kernel<<<dim3((unsigned)sp.st0[nb], XV_SLICES), 256, 0, side.s>>>(xt, V, ysw, d, ldg, pp.ktot, sp, state, i);
            pca_xty_kernel<<<dim3((unsigned)((ldg + 31) / 32), (unsigned)nb), 256, 0, side.s>>>(xt, ysw, WB, d, ldg, pp.ktot, sp, state, i);
            if ((e = cudaEventRecord(side.join, side.s)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(W, 0, (size_t)nj * ldg * 8, st)) != cudaSuccess) return e;
            pca_matvec_sym_kernel<<<dim3((unsigned)((ldg + 63) / 64), SYM_SPLIT, (unsigned)njc), 256, 0, st>>>(G, VT, W, ldg, nj, state, i);
            if ((e = cudaStreamWaitEvent(st, side.join, 0)) != cudaSuccess) return e;
            pca_update_kernel<<<nj, vt, 0, st>>>(V, VT, W, WB, stats, ldg, nj, tol * tol, state, i);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(h_stats.data(), stats, h_stats.size() * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(&h_state, state, sizeof(IterState), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        iters = h_state.iters;
        worst = 0.0;
        for (int j = 0; j < nj; ++j) worst = std::max(worst, h_stats[2 * j + 1]);
        bool conv = h_state.done != 0;
        for (int i = 0; i < batch; ++i) conv = conv || h_state.cnt[i] == nj;
        if (conv) break;
    }
    // stats[2j] is the Rayleigh quotient of the vector BEFORE the last normalisation; at convergence it
    // differs from that of V by O(tol^2).
    pca_finish_kernel<<<nj, vt, 0, st>>>(V, stats, colsum, coltot, pp, ldg, d_explained, d_sigma2, d_component);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    tm.mark();
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    tm.finish();
    if (h_iters) *h_iters = iters;
    if (h_residual) *h_residual = std::sqrt(worst);
    return cudaSuccess;
}

}  // namespace worm
