// postproc.cu -- imaging, RG blocking and bond counting of dumped worm configurations (sm_100a).
//
// All three are streaming integer kernels bounded by HBM bandwidth (no reuse beyond a 2-row
// halo), so the design rules are: 16-byte vector loads/stores along the contiguous (second)
// pixel index, every input sector read once from DRAM (neighbour reuse through L1/L2), grids of
// a few waves of 148 SMs.  No tensor cores: there is no contraction here.
//
//   image : bonds.py:199-239 (odd-parity filter) + bonds.py:241-319 (rasterise to [2L][2L])
//   block : block_configs.py:6-71 == iterated_blocking.py:6-59 (variant 0), bonds.py:333-410 (variant 1)
//   count : count_bonds.py:110-117
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

namespace {

// ---------------------------------------------------------------------------------------------
// imaging.  Image pixel (px, py), first index = x pixel, flat index px*2L + py:
//   px even, py even : site (x = px/2, y = py/2): 1 if any of its four bonds is odd
//   px odd,  py even : +x bond of site ((px-1)/2, py/2)   (the wrap bond of x = L-1 lands on px = 2L-1)
//   px even, py odd  : +y bond of site (px/2, (py-1)/2)
//   px odd,  py odd  : 0
// One block per configuration: the 2N occupation bytes are read once (coalesced) into shared
// memory as parity bytes with a padded row pitch, then each thread writes 4 pixels (int4) of a
// pixel row; consecutive threads -> consecutive py, so stores are fully coalesced.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int par_at(const uint8_t *sp, int pitch, int x, int y, int dir)
{
    return sp[y * pitch + 2 * x + dir];
}

// Fast path (vectorisable shapes): a block handles `cpb` configurations per iteration (cpb > 1 for
// small lattices, so that an iteration always moves ~64 KiB), and the occupation bytes of the NEXT
// iteration are already in flight (registers) while the current one is written out: the store
// phase hides the load latency, and the two block-wide syncs are amortised over cpb configs.
constexpr int IMG_THREADS = 256;
constexpr int IMG_MAXV = 8;             // uint4 of occupations per thread per iteration

__global__ void __launch_bounds__(IMG_THREADS)
image_kernel_pipe(int L, int lgL, int cpb, int64_t n, const uint8_t *__restrict__ bonds, int32_t *__restrict__ img)
{
    // L = 2^lgL: every index split below is a shift / mask
    extern __shared__ __align__(16) uint8_t sp[];
    const int pitch = 2 * L + 2;                 // bytes per lattice row y (x-major inside), padded
    const int tile = L * pitch;                  // parity bytes of one configuration
    const int N = L * L, nb2 = 2 * N, W = 2 * L;
    const int vpc = nb2 / 16;                    // uint4 per configuration
    const int vtot = cpb * vpc;                  // uint4 per iteration (<= IMG_THREADS * IMG_MAXV)
    const int qn = W / 4;                        // int4 per pixel row
    const int opc = W * qn;                      // int4 per image
    uint4 v[IMG_MAXV];
    auto fetch = [&](int64_t cfg0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(bonds + cfg0 * (int64_t)nb2);
        const int64_t lim = (n - cfg0) * vpc;    // uint4 left in the array
#pragma unroll
        for (int k = 0; k < IMG_MAXV; ++k) {
            const int i = threadIdx.x + k * IMG_THREADS;
            v[k] = (i < vtot && i < lim) ? s4[i] : make_uint4(0, 0, 0, 0);
        }
    };
    int64_t cfg0 = (int64_t)blockIdx.x * cpb;
    if (cfg0 < n) fetch(cfg0);
    for (; cfg0 < n; cfg0 += (int64_t)gridDim.x * cpb) {
        __syncthreads();                         // the previous iteration's reads of the tiles are done
#pragma unroll
        for (int k = 0; k < IMG_MAXV; ++k) {
            const int i = threadIdx.x + k * IMG_THREADS;
            if (i < vtot) {
                const int c = i >> (2 * lgL - 3), iv = i & (vpc - 1);           // vpc = 2^(2 lgL - 3)
                uint8_t *t = sp + c * tile;
                const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = 16 * iv + j;
                    const int site = b >> 1, y = site >> lgL, x = site & (L - 1);
                    t[y * pitch + 2 * x + (b & 1)] = (uint8_t)((w[j >> 2] >> (8 * (j & 3))) & 1u);
                }
            }
        }
        __syncthreads();
        const int64_t nxt = cfg0 + (int64_t)gridDim.x * cpb;
        if (nxt < n) fetch(nxt);                 // in flight during the store phase below
        const int64_t here = (n - cfg0 < cpb) ? (n - cfg0) : cpb;
        int4 *dst = reinterpret_cast<int4 *>(img + cfg0 * (int64_t)W * W);
        for (int t = threadIdx.x; t < (int)here * opc; t += IMG_THREADS) {
            const int c = t >> (2 * lgL), r = t & (opc - 1);                  // opc = 2L * L/2 = 2^(2 lgL)
            const uint8_t *tp = sp + c * tile;
            const int px = r >> (lgL - 1), qq = r & (qn - 1);                // qn = L/2
            const int x = px >> 1, y0 = 2 * qq;        // py = 4*qq .. 4*qq+3  -> y0, y0+1
            int4 o;
            if (px & 1) {
                o.x = par_at(tp, pitch, x, y0, 0);
                o.y = 0;
                o.z = par_at(tp, pitch, x, y0 + 1, 0);
                o.w = 0;
            } else {
                const int xm = (x == 0) ? L - 1 : x - 1;
                const int ym = (y0 == 0) ? L - 1 : y0 - 1;
                const int yb0 = par_at(tp, pitch, x, y0, 1);
                const int yb1 = par_at(tp, pitch, x, y0 + 1, 1);
                o.x = par_at(tp, pitch, x, y0, 0) | yb0 | par_at(tp, pitch, xm, y0, 0) | par_at(tp, pitch, x, ym, 1);
                o.y = yb0;
                o.z = par_at(tp, pitch, x, y0 + 1, 0) | yb1 | par_at(tp, pitch, xm, y0 + 1, 0) | yb0;
                o.w = yb1;
            }
            dst[t] = o;
        }
    }
}

// Generic path: any L, any alignment; one configuration per block iteration.
__global__ void __launch_bounds__(256)
image_kernel(int L, int64_t n, const uint8_t *__restrict__ bonds, int32_t *__restrict__ img, int vec_out)
{
    extern __shared__ __align__(16) uint8_t sp[];
    const int pitch = 2 * L + 2;                 // bytes per lattice row y (x-major inside), padded
    const int N = L * L, nb2 = 2 * N, W = 2 * L;
    for (int64_t cfg = blockIdx.x; cfg < n; cfg += gridDim.x) {
        const uint8_t *src = bonds + cfg * (int64_t)nb2;
        __syncthreads();
        for (int b = threadIdx.x; b < nb2; b += blockDim.x) {
            const int site = b >> 1, y = site / L, x = site - y * L;
            sp[y * pitch + 2 * x + (b & 1)] = src[b] & 1u;
        }
        __syncthreads();
        int32_t *dst = img + cfg * (int64_t)W * W;
        for (int t = threadIdx.x; t < W * W; t += blockDim.x) {
            const int px = t / W, py = t - px * W;
            const int x = px >> 1, y = py >> 1;
            int v;
            if ((px & 1) && (py & 1)) v = 0;
            else if (px & 1) v = par_at(sp, pitch, x, y, 0);
            else if (py & 1) v = par_at(sp, pitch, x, y, 1);
            else {
                const int xm = (x == 0) ? L - 1 : x - 1;
                const int ym = (y == 0) ? L - 1 : y - 1;
                v = par_at(sp, pitch, x, y, 0) | par_at(sp, pitch, x, y, 1) | par_at(sp, pitch, xm, y, 0) |
                    par_at(sp, pitch, x, ym, 1);
            }
            dst[t] = v;
        }
        (void)vec_out;
    }
}

// ---------------------------------------------------------------------------------------------
// blocking.  A [W][W] -> B [Wo][Wo], Wo = W/2, blocks (p, q), p, q < h = Wo/2:
//   xb(p,q) = f(A[4p][4q+3], A[4p+2][4q+3])  -> B[2p][2q+1]        (block_configs.py:35,38,58)
//   yb(p,q) = f(A[4p+3][4q], A[4p+3][4q+2])  -> B[2p+1][2q]        (:36,39,59)
//   B[2p][2q] = site(xb, yb) then forced to 1 if B[2p][2q-1] or B[2p-1][2q] is non-zero, indices
//   taken mod Wo like Python's negative indexing (:56-57,61-65); B[odd][odd] = 0.
//   f = xor for double_bonds_value 0; otherwise (a+b == 2 ? v : a+b == 1 ? 1 : 0)  (:41-55).
// For odd Wo the wrapped neighbour is a pixel no block ever writes, i.e. 0.
// variant 1 (bonds.py:381-409): xor first, then pairs equal to [1,1] overwrite with block_val;
// the fix-up sets 1 if a neighbour == 1 and then block_val if a neighbour == block_val.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int comb(int a, int b, int v, int variant)
{
    if (variant == 0) {
        if (v == 0) return a ^ b;
        const int s = a + b;
        return s == 2 ? v : (s == 1 ? 1 : 0);
    }
    int r = a ^ b;
    if (v != 0 && a == 1 && b == 1) r = v;
    return r;
}

__device__ __forceinline__ int site_of(int xb, int yb, int ex0, int ex1, int ey0, int ey1, int v, int variant)
{
    if (variant == 0) return xb ? xb : yb;                 // Python `xb or yb`
    int s = (ex0 ^ ex1) ? (ex0 ^ ex1) : (ey0 ^ ey1);       // bonds.py:388
    if (v != 0) {
        if (ex0 == 1 && ex1 == 1) s = v;                   // :393-395
        if (ey0 == 1 && ey1 == 1) s = v;                   // :396-398
    }
    return s;
}

__device__ __forceinline__ int fixup(int s, int left, int down, int v, int variant)
{
    if (variant == 0) return (left || down) ? 1 : s;       // block_configs.py:64-65
    if (left == 1 || down == 1) s = 1;                     // bonds.py:405-406
    if (v != 0 && (left == v || down == v)) s = v;         // :407-409
    return s;
}

// generic scalar kernel: one thread per block (p, q); any W
__global__ void __launch_bounds__(256)
block_kernel_scalar(int W, int64_t n, const int32_t *__restrict__ in, int32_t *__restrict__ out, int v, int variant)
{
    const int Wo = W / 2, h = Wo / 2;
    const int64_t per = (int64_t)Wo * Wo;
    const int64_t total = n * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cfg = t / per;
        const int rem = (int)(t - cfg * per);
        const int i = rem / Wo, j = rem - i * Wo;
        const int32_t *A = in + cfg * (int64_t)W * W;
        int val = 0;
        const int p = i >> 1, q = j >> 1;
        if (p < h && q < h) {
            auto XB = [&](int pp, int qq, int &e0, int &e1) {
                e0 = A[(4 * pp) * W + 4 * qq + 3]; e1 = A[(4 * pp + 2) * W + 4 * qq + 3];
                return comb(e0, e1, v, variant);
            };
            auto YB = [&](int pp, int qq, int &e0, int &e1) {
                e0 = A[(4 * pp + 3) * W + 4 * qq]; e1 = A[(4 * pp + 3) * W + 4 * qq + 2];
                return comb(e0, e1, v, variant);
            };
            int a0, a1, b0, b1;
            if ((i & 1) == 0 && (j & 1) == 1) val = XB(p, q, a0, a1);
            else if ((i & 1) == 1 && (j & 1) == 0) val = YB(p, q, b0, b1);
            else if ((i & 1) == 0 && (j & 1) == 0) {
                const int xb = XB(p, q, a0, a1), yb = YB(p, q, b0, b1);
                int s = site_of(xb, yb, a0, a1, b0, b1, v, variant);
                int left = 0, down = 0, t0, t1;
                // B[2p][2q-1]: q > 0 -> xb(p, q-1); q == 0 -> column Wo-1: a bond column iff Wo even
                if (q > 0) left = XB(p, q - 1, t0, t1);
                else if ((Wo & 1) == 0) left = XB(p, h - 1, t0, t1);
                if (p > 0) down = YB(p - 1, q, t0, t1);
                else if ((Wo & 1) == 0) down = YB(h - 1, q, t0, t1);
                val = fixup(s, left, down, v, variant);
            }
        }
        out[t] = val;
    }
}

// vector kernel (Wo % 4 == 0, hence W % 8 == 0): one thread per (config, block row p, 4 output
// columns) produces B rows 2p and 2p+1 with two int4 stores from ten int4 loads; consecutive
// threads -> consecutive columns (coalesced); the halo loads hit L1.
__global__ void __launch_bounds__(256)
block_kernel_vec(int W, int64_t n, const int32_t *__restrict__ in, int32_t *__restrict__ out, int v, int variant)
{
    const int Wo = W / 2, h = Wo / 2;
    const int tq = Wo / 4;                        // threads per block row
    const int64_t per = (int64_t)h * tq;
    const int64_t total = n * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cfg = t / per;
        const int rem = (int)(t - cfg * per);
        const int p = rem / tq, tt = rem - p * tq;           // output columns 4tt..4tt+3 = blocks q0 = 2tt, 2tt+1
        const int32_t *A = in + cfg * (int64_t)W * W;
        const int4 *r0 = reinterpret_cast<const int4 *>(A + (int64_t)(4 * p) * W);
        const int4 *r2 = reinterpret_cast<const int4 *>(A + (int64_t)(4 * p + 2) * W);
        const int4 *r3 = reinterpret_cast<const int4 *>(A + (int64_t)(4 * p + 3) * W);
        const int pm = (p == 0) ? h - 1 : p - 1;
        const int4 *rm = reinterpret_cast<const int4 *>(A + (int64_t)(4 * pm + 3) * W);
        const int c0 = 2 * tt, c1 = 2 * tt + 1;               // int4 index of columns 8tt.. and 8tt+4..
        const int cm = (tt == 0) ? 2 * tq - 1 : c0 - 1;       // int4 holding column 8tt-1 (wraps to W-1)
        const int4 a0 = r0[c0], a1 = r0[c1], am = r0[cm];
        const int4 b0 = r2[c0], b1 = r2[c1], bm = r2[cm];
        const int4 y0 = r3[c0], y1 = r3[c1];
        const int4 d0 = rm[c0], d1 = rm[c1];
        const int xb0 = comb(a0.w, b0.w, v, variant), xb1 = comb(a1.w, b1.w, v, variant);
        const int xbm = comb(am.w, bm.w, v, variant);
        const int yb0 = comb(y0.x, y0.z, v, variant), yb1 = comb(y1.x, y1.z, v, variant);
        const int dn0 = comb(d0.x, d0.z, v, variant), dn1 = comb(d1.x, d1.z, v, variant);
        int s0 = site_of(xb0, yb0, a0.w, b0.w, y0.x, y0.z, v, variant);
        int s1 = site_of(xb1, yb1, a1.w, b1.w, y1.x, y1.z, v, variant);
        s0 = fixup(s0, xbm, dn0, v, variant);
        s1 = fixup(s1, xb0, dn1, v, variant);
        int32_t *B = out + cfg * (int64_t)Wo * Wo;
        reinterpret_cast<int4 *>(B + (int64_t)(2 * p) * Wo)[tt] = make_int4(s0, xb0, s1, xb1);
        reinterpret_cast<int4 *>(B + (int64_t)(2 * p + 1) * Wo)[tt] = make_int4(yb0, 0, yb1, 0);
    }
}

// ---------------------------------------------------------------------------------------------
// counting: n = sum over (i + j) odd.  One warp per configuration (grid-stride), int4 loads when
// W % 4 == 0: in an even row the odd columns are .y/.w, in an odd row .x/.z.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_kernel(int W, int64_t n, const int32_t *__restrict__ img, int64_t *__restrict__ cnt, int vec)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int per = W * W;
    for (int64_t cfg = warp; cfg < n; cfg += nwarps) {
        const int32_t *A = img + cfg * (int64_t)per;
        long long s = 0;
        if (vec) {
            const int4 *A4 = reinterpret_cast<const int4 *>(A);
            const int qn = W / 4, tot = per / 4;
            for (int t = lane; t < tot; t += 32) {
                const int4 vv = A4[t];
                const int i = t / qn;
                s += (i & 1) ? ((long long)vv.x + vv.z) : ((long long)vv.y + vv.w);
            }
        } else {
            for (int t = lane; t < per; t += 32) {
                const int i = t / W, j = t - i * W;
                if ((i + j) & 1) s += A[t];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) cnt[cfg] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused C3 pipeline: dumped occupations -> parity bits -> (images) -> every "1+1=0" blocking level -> counts,
// one launch (bonds.py:199-319 -> iterated_blocking.py:6-59 x k -> count_bonds.py:107-123).
//
// Everything the chain computes is a function of the odd-parity bonds.  With X(x, y) / Y(x, y) the parities
// of the +x / +y bond of site (x, y), one blocking step of a bond image is, in bond language
// (block_configs.py:35-39,58-59 read through the pixel map of bonds.py:241-319):
//     Y'(p, q) = Y(2p, 2q+1) ^ Y(2p+1, 2q+1)        (the two +y bonds leaving the 2x2 block; pixel B[2p][2q+1])
//     X'(p, q) = X(2p+1, 2q) ^ X(2p+1, 2q+1)        (the two +x bonds leaving it;             pixel B[2p+1][2q])
// the site pixel is the OR of the site's four incident bonds at every level (imaging: bonds.py:271-299;
// blocking: :56-57 + the fix-up :61-65), and the bond count (count_bonds.py:110-117, pixels with i + j odd)
// is popcount(X) + popcount(Y).  So a configuration is 2 L^2 BITS: a warp reads its 2 L^2 occupation bytes
// once (coalesced 16-byte loads), keeps the bit rows of every level in shared memory (L^2 / 4 bytes at
// level 0) and writes `levels` counts -- 2 L^2 bytes of HBM traffic per configuration instead of the
// 16 L^2-byte int32 image and its pyramid.  Images of any level are written only on request.
// Power-of-two L >= 4; level k has lattice size L >> k (>= 2).
// ---------------------------------------------------------------------------------------------
constexpr int PIPE_WARPS = 8;

__device__ __forceinline__ uint32_t squeeze_even(uint32_t w)      // bits 0, 2, 4, ... -> bits 0, 1, 2, ... (16 of them)
{
    w &= 0x55555555u;
    w = (w | (w >> 1)) & 0x33333333u;
    w = (w | (w >> 2)) & 0x0f0f0f0fu;
    w = (w | (w >> 4)) & 0x00ff00ffu;
    return (w | (w >> 8)) & 0x0000ffffu;
}

struct PipeImages { int32_t *level[12]; };

// bit rows of one level: X rows then Y rows, `wpr` words per row (row y of X at X[y * wpr ...])
__device__ __forceinline__ void pipe_write_image(const uint32_t *X, const uint32_t *Y, int Ll, int wpr, int32_t *img, int lane)
{
    const int W = 2 * Ll, qn = W / 4;              // int4 per pixel row (W >= 4)
    auto xb = [&](int x, int y) { return (int)((X[y * wpr + (x >> 5)] >> (x & 31)) & 1u); };
    auto yb = [&](int x, int y) { return (int)((Y[y * wpr + (x >> 5)] >> (x & 31)) & 1u); };
    int4 *dst = reinterpret_cast<int4 *>(img);
    for (int t = lane; t < W * qn; t += 32) {
        const int px = t / qn, qq = t - px * qn;
        const int x = px >> 1, y0 = 2 * qq;
        int4 o;
        if (px & 1) {
            o.x = xb(x, y0); o.y = 0;
            o.z = (y0 + 1 < Ll) ? xb(x, y0 + 1) : 0; o.w = 0;
        } else {
            const int xm = x ? x - 1 : Ll - 1, ym = y0 ? y0 - 1 : Ll - 1;
            const int yb0 = yb(x, y0), yb1 = yb(x, y0 + 1);
            o.x = xb(x, y0) | yb0 | xb(xm, y0) | yb(x, ym);
            o.y = yb0;
            o.z = xb(x, y0 + 1) | yb1 | xb(xm, y0 + 1) | yb0;
            o.w = yb1;
        }
        dst[t] = o;
    }
}

__global__ void __launch_bounds__(PIPE_WARPS * 32)
pipeline_kernel(int L, int64_t n, const uint8_t *__restrict__ bonds, int levels, int64_t *__restrict__ counts,
                const PipeImages imgs, int words_per_warp)
{
    extern __shared__ __align__(16) uint32_t pipe_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *buf = pipe_sm + (size_t)warp * words_per_warp;
    const int N = L * L, nb2 = 2 * N;
    const int wpr0 = L >= 32 ? L / 32 : 1;
    const int lvl0_words = L * wpr0;                         // words of the X (or Y) rows at level 0
    // level 0 lives in [0, 2 * lvl0_words), level 1 behind it, level 2 back at 0 (it is at most 1/4 the size) ...
    for (int64_t cfg = (int64_t)blockIdx.x * PIPE_WARPS + warp; cfg < n; cfg += (int64_t)gridDim.x * PIPE_WARPS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bonds + cfg * (int64_t)nb2);
        uint32_t *X = buf, *Y = buf + lvl0_words;
        uint8_t *Xb = reinterpret_cast<uint8_t *>(X), *Yb = reinterpret_cast<uint8_t *>(Y);
        if (L < 32) {
            for (int i = lane; i < 2 * lvl0_words; i += 32) buf[i] = 0u;      // rows are padded to a word
            __syncwarp();
        }
        // 16 occupation bytes = 8 sites = 8 X parities (bytes 0, 2 of each word) + 8 Y parities (bytes 1, 3).
        // Each word's two X bits sit at bits 0 / 16 and its Y bits at bits 8 / 24: the four words are merged with
        // shifts of 2 (one IMAD each), then one fold by 15 brings the odd-site bits next to the even-site ones.
        for (int i = lane; i < nb2 / 16; i += 32) {
            const uint4 v = src[i];
            const uint32_t a0 = v.x & 0x01010101u, a1 = v.y & 0x01010101u, a2 = v.z & 0x01010101u, a3 = v.w & 0x01010101u;
            const uint32_t m = a0 + a1 * 4u + a2 * 16u + a3 * 64u;            // bits 2k (+16): X of site pair k; bits 2k + 8 (+16): Y
            const uint32_t f = m | (m >> 15);                                   // bit 2k + 1 <- bit 2k + 16 (the odd site of the pair)
            const uint32_t xs = f & 0xffu, ys = (f >> 8) & 0xffu;
            if (L >= 8) {
                const int s0 = 8 * i, y = s0 / L, x0 = s0 - y * L;
                Xb[y * wpr0 * 4 + (x0 >> 3)] = (uint8_t)xs;
                Yb[y * wpr0 * 4 + (x0 >> 3)] = (uint8_t)ys;
            } else {                                                          // L = 4: two rows per piece
                Xb[(2 * i) * 4] = (uint8_t)(xs & 15u); Xb[(2 * i + 1) * 4] = (uint8_t)(xs >> 4);
                Yb[(2 * i) * 4] = (uint8_t)(ys & 15u); Yb[(2 * i + 1) * 4] = (uint8_t)(ys >> 4);
            }
        }
        __syncwarp();
        int Ll = L, wpr = wpr0;
        for (int lv = 0; lv < levels; ++lv) {
            int c = 0;
            for (int i = lane; i < Ll * wpr; i += 32) c += __popc(X[i]) + __popc(Y[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && counts) counts[cfg * levels + lv] = c;
            if (imgs.level[lv]) pipe_write_image(X, Y, Ll, wpr, imgs.level[lv] + cfg * (int64_t)(4 * Ll * Ll), lane);
            if (lv + 1 == levels) break;
            // one blocking step into the other region
            const int Ln = Ll / 2, wprn = Ln >= 32 ? Ln / 32 : 1;
            uint32_t *Xn = (X == buf) ? buf + 2 * lvl0_words : buf;
            uint32_t *Yn = Xn + Ln * wprn;
            for (int i = lane; i < Ln * wprn; i += 32) {
                const int q = i / wprn, j = i - q * wprn;
                // X'(p, q), p = 32 j ..: odd bits of X rows 2q, 2q+1 (xor), words 2j and 2j+1
                uint32_t t0 = X[(2 * q) * wpr + 2 * j] ^ X[(2 * q + 1) * wpr + 2 * j];
                uint32_t u0 = Y[(2 * q + 1) * wpr + 2 * j];
                uint32_t xn = squeeze_even(t0 >> 1);
                uint32_t yn = squeeze_even(u0 ^ (u0 >> 1));
                if (2 * j + 1 < wpr) {
                    const uint32_t t1 = X[(2 * q) * wpr + 2 * j + 1] ^ X[(2 * q + 1) * wpr + 2 * j + 1];
                    const uint32_t u1 = Y[(2 * q + 1) * wpr + 2 * j + 1];
                    xn |= squeeze_even(t1 >> 1) << 16;
                    yn |= squeeze_even(u1 ^ (u1 >> 1)) << 16;
                }
                Xn[i] = xn; Yn[i] = yn;
            }
            __syncwarp();
            X = Xn; Y = Yn; Ll = Ln; wpr = wprn;
        }
        __syncwarp();
    }
}

int grid_for(int64_t work_items, int threads, int waves_cap = 148 * 16)
{
    int64_t b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > waves_cap) b = waves_cap;
    return (int)b;
}

}  // namespace

cudaError_t launch_pipeline(int L, int64_t n, const uint8_t *d_bonds, int levels, int64_t *d_counts,
                            int32_t *const *h_level_images, int smem_optin, cudaStream_t st);

cudaError_t launch_image(int L, int64_t n, const uint8_t *d_bonds, int32_t *d_img, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    // small power-of-two lattices: the fused kernel's image writer (parity bits in shared memory, pixels expanded from
    // bit rows) beats the byte-transposing kernel below (L = 32: 0.82 against 0.75 of the HBM peak)
    if (L >= 4 && L <= 32 && !(L & (L - 1)) && (((uintptr_t)d_img & 15) == 0) && (((uintptr_t)d_bonds & 15) == 0)) {
        int32_t *lv[1] = {d_img};
        return launch_pipeline(L, n, d_bonds, 1, nullptr, lv, 200 * 1024, st);
    }
    const int nb2 = 2 * L * L;
    const size_t tile = (size_t)L * (2 * L + 2);
    const bool aligned = (nb2 % 16 == 0) && ((2 * L) % 4 == 0) && (((uintptr_t)d_img & 15) == 0) &&
                         (((uintptr_t)d_bonds & 15) == 0);
    const int vpc = nb2 / 16;
    const bool pow2 = L >= 4 && (L & (L - 1)) == 0;
    int lgL = 0;
    while ((1 << lgL) < L) ++lgL;
    if (aligned && pow2 && vpc <= IMG_THREADS * IMG_MAXV) {
        // configurations per block iteration: ~64 KiB of output, within the register prefetch budget
        int cpb = (16384 / (4 * L * L)) > 0 ? 16384 / (4 * L * L) : 1;
        if (cpb * vpc > IMG_THREADS * IMG_MAXV) cpb = IMG_THREADS * IMG_MAXV / vpc;
        if ((int64_t)cpb > n) cpb = (int)n;
        const size_t smem = (size_t)cpb * tile + 16;
        cudaError_t e = cudaFuncSetAttribute(image_kernel_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t iters = (n + cpb - 1) / cpb;
        const int grid = (int)(iters < 148 * 8 ? iters : 148 * 8);
        image_kernel_pipe<<<grid, IMG_THREADS, smem, st>>>(L, lgL, cpb, n, d_bonds, d_img);
        return cudaGetLastError();
    }
    const size_t smem = tile + 16;
    cudaError_t e = cudaFuncSetAttribute(image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)(n < 148 * 8 ? n : 148 * 8);
    image_kernel<<<grid, 256, smem, st>>>(L, n, d_bonds, d_img, 0);
    return cudaGetLastError();
}

cudaError_t launch_block(int W, int64_t n, const int32_t *d_in, int32_t *d_out, int dbv, int variant, cudaStream_t st)
{
    if (n == 0 || W < 2) return cudaSuccess;
    const int Wo = W / 2;
    if ((Wo & 3) == 0 && ((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0) {
        const int64_t items = n * (int64_t)(Wo / 2) * (Wo / 4);
        block_kernel_vec<<<grid_for(items, 256), 256, 0, st>>>(W, n, d_in, d_out, dbv, variant);
    } else {
        const int64_t items = n * (int64_t)Wo * Wo;
        block_kernel_scalar<<<grid_for(items, 256), 256, 0, st>>>(W, n, d_in, d_out, dbv, variant);
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Boundaries between the black and white regions of thresholded grey images (image_analysis/image.py:46-80)
// ---------------------------------------------------------------------------------------------------
// bw(x, y) = image(x, y) >= cutoff.  link0(x, y) = bw(x, y) ^ bw(x, y-1), link1(x, y) = bw(x, y) ^ bw(x-1, y)
// (periodic).  Output [2Nx][2Ny]: (2x+1, 2y) = link0, (2x, 2y+1) = link1, (2x, 2y) = any of link0(x, y),
// link1(x, y), link0(x-1, y), link1(x, y-1); (2x+1, 2y+1) = 0.  One thread per site, one launch for all
// images x cutoffs; each thread writes two 8-byte pairs, consecutive threads consecutive y.
__global__ void __launch_bounds__(256) boundaries_kernel(int Nx, int Ny, int64_t n, int m, const double *__restrict__ img,
                                                         const double *__restrict__ cutoffs, int32_t *__restrict__ out)
{
    // generic shape: one thread per site; blockIdx.x = image * m + cutoff, blockIdx.y = block of sites
    const int sites = Nx * Ny;
    const int64_t im = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= sites) return;
    const int x = s / Ny, y = s - x * Ny;
    const double cut = cutoffs[im % m];
    const double *src = img + (im / m) * sites;
    const int xm = x ? x - 1 : Nx - 1, ym = y ? y - 1 : Ny - 1;
    const int b = src[(size_t)x * Ny + y] >= cut, bxm = src[(size_t)xm * Ny + y] >= cut;
    const int bym = src[(size_t)x * Ny + ym] >= cut, bxy = src[(size_t)xm * Ny + ym] >= cut;
    const int l0 = b ^ bym, l1 = b ^ bxm;
    const int l0_up = bxm ^ bxy, l1_left = bym ^ bxy;
    int32_t *dst = out + im * 4 * sites + (size_t)(2 * x) * (2 * Ny) + 2 * y;
    *reinterpret_cast<int2 *>(dst) = make_int2((l0 | l1 | l0_up | l1_left) ? 1 : 0, l1);
    *reinterpret_cast<int2 *>(dst + 2 * Ny) = make_int2(l0, 0);
}

// Even Ny: one thread per pair of sites (x, 2k), (x, 2k+1): two 16-byte stores, six pixel reads.
__global__ void __launch_bounds__(256) boundaries_kernel_pair(int Nx, int Ny, int64_t n, int m, const double *__restrict__ img,
                                                              const double *__restrict__ cutoffs, int32_t *__restrict__ out)
{
    const int hy = Ny >> 1, pairs = Nx * hy;
    const int64_t im = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= pairs) return;
    const int x = s / hy, y = 2 * (s - x * hy);
    const double cut = cutoffs[im % m];
    const double *row = img + (im / m) * (size_t)Nx * Ny + (size_t)x * Ny;
    const double *rowm = img + (im / m) * (size_t)Nx * Ny + (size_t)(x ? x - 1 : Nx - 1) * Ny;
    const int ym = y ? y - 1 : Ny - 1;
    const double2 c = *reinterpret_cast<const double2 *>(row + y), u = *reinterpret_cast<const double2 *>(rowm + y);
    const int b0 = c.x >= cut, b1 = c.y >= cut, u0 = u.x >= cut, u1 = u.y >= cut;
    const int bl = row[ym] >= cut, ul = rowm[ym] >= cut;     // left neighbours of the pair
    // site (x, y): link0 = b0 ^ bl, link1 = b0 ^ u0, link0 of (x-1, y) = u0 ^ ul, link1 of (x, y-1) = bl ^ ul
    const int l0a = b0 ^ bl, l1a = b0 ^ u0, sa = l0a | l1a | (u0 ^ ul) | (bl ^ ul);
    // site (x, y+1): left neighbour is (x, y)
    const int l0b = b1 ^ b0, l1b = b1 ^ u1, sb = l0b | l1b | (u1 ^ u0) | (b0 ^ u0);
    int32_t *dst = out + im * 4 * (int64_t)Nx * Ny + (size_t)(2 * x) * (2 * Ny) + 2 * y;
    *reinterpret_cast<int4 *>(dst) = make_int4(sa, l1a, sb, l1b);
    *reinterpret_cast<int4 *>(dst + 2 * Ny) = make_int4(l0a, 0, l0b, 0);
}

// The same with the loop over the cutoffs INSIDE the thread: the six pixels of a site pair are read once and
// compared against every cutoff, so per cutoff a thread only compares and stores (the version above re-reads the
// pixels and redoes the index arithmetic for every cutoff: 68 % of its issue slots, 0.70-0.77 of the HBM peak).
// blockIdx.x = image, blockIdx.y = block of site pairs.
__global__ void __launch_bounds__(256) boundaries_kernel_allcuts(int Nx, int Ny, int64_t n, int m, const double *__restrict__ img,
                                                                 const double *__restrict__ cutoffs, int32_t *__restrict__ out)
{
    const int hy = Ny >> 1, pairs = Nx * hy;
    const int64_t image = blockIdx.x;
    const int s = blockIdx.y * blockDim.x + threadIdx.x;
    if (s >= pairs) return;
    const int x = s / hy, y = 2 * (s - x * hy);
    const double *row = img + image * (size_t)Nx * Ny + (size_t)x * Ny;
    const double *rowm = img + image * (size_t)Nx * Ny + (size_t)(x ? x - 1 : Nx - 1) * Ny;
    const int ym = y ? y - 1 : Ny - 1;
    const double2 c = *reinterpret_cast<const double2 *>(row + y), u = *reinterpret_cast<const double2 *>(rowm + y);
    const double pl = row[ym], pul = rowm[ym];
    const size_t per = 4 * (size_t)Nx * Ny;
    int32_t *dst = out + image * m * per + (size_t)(2 * x) * (2 * Ny) + 2 * y;
    for (int k = 0; k < m; ++k, dst += per) {
        const double cut = cutoffs[k];
        const int b0 = c.x >= cut, b1 = c.y >= cut, u0 = u.x >= cut, u1 = u.y >= cut;
        const int bl = pl >= cut, ul = pul >= cut;
        const int l0a = b0 ^ bl, l1a = b0 ^ u0, sa = l0a | l1a | (u0 ^ ul) | (bl ^ ul);
        const int l0b = b1 ^ b0, l1b = b1 ^ u1, sb = l0b | l1b | (u1 ^ u0) | (b0 ^ u0);
        *reinterpret_cast<int4 *>(dst) = make_int4(sa, l1a, sb, l1b);
        *reinterpret_cast<int4 *>(dst + 2 * Ny) = make_int4(l0a, 0, l0b, 0);
    }
}

cudaError_t launch_boundaries(int Nx, int Ny, int64_t n, int m, const double *d_img, const double *d_cutoffs, int32_t *d_out,
                              cudaStream_t st)
{
    const int64_t ims = n * m;
    if (ims == 0) return cudaSuccess;
    if (ims > 0x7fffffffll) return cudaErrorInvalidValue;
    const bool pair = (Ny & 1) == 0 && ((uintptr_t)d_img & 15) == 0 && ((uintptr_t)d_out & 15) == 0;
    const int items = pair ? Nx * (Ny / 2) : Nx * Ny;
    const int nblk = (items + 255) / 256;
    const int threads = ((items + nblk - 1) / nblk + 31) / 32 * 32;   // even split of an image over its blocks
    if (pair && m >= 2) {
        const dim3 grid_a((unsigned)n, (unsigned)nblk);
        boundaries_kernel_allcuts<<<grid_a, threads, 0, st>>>(Nx, Ny, n, m, d_img, d_cutoffs, d_out);
        return cudaGetLastError();
    }
    const dim3 grid((unsigned)ims, (unsigned)nblk);
    if (pair) boundaries_kernel_pair<<<grid, threads, 0, st>>>(Nx, Ny, n, m, d_img, d_cutoffs, d_out);
    else boundaries_kernel<<<grid, threads, 0, st>>>(Nx, Ny, n, m, d_img, d_cutoffs, d_out);
    return cudaGetLastError();
}

cudaError_t launch_pipeline(int L, int64_t n, const uint8_t *d_bonds, int levels, int64_t *d_counts,
                            int32_t *const *h_level_images, int smem_optin, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    PipeImages imgs{};
    for (int k = 0; k < 12; ++k) imgs.level[k] = (h_level_images && k < levels) ? h_level_images[k] : nullptr;
    const int wpr0 = L >= 32 ? L / 32 : 1;
    const int lvl0 = L * wpr0;
    const int Ln = L / 2, wprn = Ln >= 32 ? Ln / 32 : 1;
    const int words_per_warp = 2 * lvl0 + 2 * Ln * wprn;
    const size_t smem = (size_t)PIPE_WARPS * words_per_warp * 4;
    if (smem > (size_t)smem_optin) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(pipeline_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t blocks_needed = (n + PIPE_WARPS - 1) / PIPE_WARPS;
    // resident blocks per SM by shared memory (<= 8 by threads), a few waves at most
    int per_sm = (int)((size_t)smem_optin / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)148 * per_sm * 4;
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    pipeline_kernel<<<grid, PIPE_WARPS * 32, smem, st>>>(L, n, d_bonds, levels, d_counts, imgs, words_per_warp);
    return cudaGetLastError();
}

cudaError_t launch_count(int W, int64_t n, const int32_t *d_img, int64_t *d_count, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const int64_t threads_needed = n * 32;
    const int vec = (W % 4 == 0) && (((uintptr_t)d_img & 15) == 0);
    count_kernel<<<grid_for(threads_needed, 256), 256, 0, st>>>(W, n, d_img, d_count, vec);
    return cudaGetLastError();
}

}  // namespace worm
