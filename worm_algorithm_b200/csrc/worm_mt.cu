// worm_mt.cu -- deterministic mode: the reference's Markov chain on the reference's random stream.
//
// One WARP per chain.  Every lane walks the loop of worm_ising_2d.cpp:111-157 redundantly (uniform
// control flow; shared-memory reads broadcast, all lanes store the same value), so the draw order
// (kick :128, direction :137, +/- coin :141, acceptance :151) and the stop rule (:133-134) are
// literally the reference's, while the parts that parallelise use the 32 lanes:
//   - MT19937 (mt19937ar.c:68-81, :113-148): the 624-word state lives in shared memory; a refill
//     twists it in three dependency-free phases ([0,227), [227,454), [454,624)) and tempers all 624
//     words in parallel, so a draw on the walk is one shared-memory load;
//   - dumps (:120-124, :160-167) are copied by the whole warp.
// The acceptance test `genrand() < P` is evaluated in integers: genrand() = u32 * 2^-32 exactly, so
// u32 * 2^-32 < P  <=>  u32 < ceil(P * 2^32), and the thresholds for P = K/(nb+1.0) and nb*inv_K
// (:143, :148; IEEE doubles, computed with the reference's own expressions on the device with
// -fmad=false) are tabulated per chain for nb < 64 at kernel start; beyond that the double
// expression itself is evaluated.  Bonds are uint8 in shared memory when 2N bytes fit, else in the
// caller's global array.
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr int THR_N = 64;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far)
{
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// In-place twist of s[624] by the 32 lanes + tempered copy in t[624].  The sequential algorithm
// updates i = 0..623 in order with s[i+397 mod 624] already NEW for i >= 227 and s[0] NEW for
// i = 623; the three phases below reproduce exactly those dependencies.
__device__ __forceinline__ void mt_refill(uint32_t *s, uint32_t *t, int lane)
{
    // phase 1: i in [0, 227): old s[i], old s[i+1], old s[i+397]
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane + 32 * k;
        if (i < 227) v[k] = mt_twist(s[i], s[i + 1], s[i + MT_M]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = lane + 32 * k;
        if (i < 227) s[i] = v[k];
    }
    __syncwarp();
    // phase 2: i in [227, 454): old s[i], old s[i+1], NEW s[i-227]
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = 227 + lane + 32 * k;
        if (i < 454) v[k] = mt_twist(s[i], s[i + 1], s[i - 227]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = 227 + lane + 32 * k;
        if (i < 454) s[i] = v[k];
    }
    __syncwarp();
    // phase 3: i in [454, 623): old s[i], old s[i+1], NEW s[i-227]; i = 623 uses NEW s[0]
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int i = 454 + lane + 32 * k;
        if (i < 623) v[k] = mt_twist(s[i], s[i + 1], s[i - 227]);
        else if (i == 623) v[k] = mt_twist(s[623], s[0], s[396]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int i = 454 + lane + 32 * k;
        if (i < MT_N) s[i] = v[k];
    }
    __syncwarp();
    for (int i = lane; i < MT_N; i += 32) t[i] = mt_temper(s[i]);
    __syncwarp();
}

// ceil(P * 2^32) clamped to [0, 2^32]: genrand() < P  <=>  u32 < this
__device__ __forceinline__ unsigned long long thr_of(double P)
{
    const double sc = ceil(ldexp(P, 32));
    if (!(sc > 0.0)) return 0ull;
    if (sc >= 4294967296.0) return 4294967296ull;
    return (unsigned long long)sc;
}

template <bool SMEM>
__global__ void __launch_bounds__(32)
worm_mt_kernel(int L, int64_t n_chains, const MtChain *__restrict__ prm,
               uint64_t num_steps, uint64_t therm, int bond_flag,
               uint8_t *__restrict__ bonds_all, int64_t *__restrict__ out,
               int32_t max_events, int32_t *__restrict__ n_events, int64_t *__restrict__ events,
               uint8_t *__restrict__ dumps,
               int32_t max_trace, int32_t *__restrict__ n_trace, int64_t *__restrict__ trace)
{
    extern __shared__ __align__(16) uint8_t smem8[];
    const int lane = threadIdx.x;
    const int64_t c = blockIdx.x;
    if (c >= n_chains) return;
    const int N = L * L;
    const int nb2 = 2 * N;
    const MtChain p = prm[c];
    uint32_t *s = reinterpret_cast<uint32_t *>(smem8);                        // MT state
    uint32_t *t = s + MT_N;                                                   // tempered outputs
    unsigned long long *thr_inc = reinterpret_cast<unsigned long long *>(t + MT_N);
    unsigned long long *thr_dec = thr_inc + THR_N;
    uint8_t *gb = bonds_all + c * (int64_t)nb2;
    uint8_t *bonds = SMEM ? reinterpret_cast<uint8_t *>(thr_dec + THR_N) : gb;

    for (int b = lane; b < nb2; b += 32) bonds[b] = 0;                        // worm_ising_2d.cpp:97-101
    for (int k = lane; k < THR_N; k += 32) {
        thr_inc[k] = thr_of(p.K / ((double)k + 1.0));                         // :143
        thr_dec[k] = thr_of((double)k * p.inv_K);                             // :148
    }
    if (lane == 0) {                                                          // init_genrand, mt19937ar.c:68-81
        uint32_t v = p.seed;
        s[0] = v;
        for (int i = 1; i < MT_N; ++i) {
            v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
            s[i] = v;
        }
    }
    __syncwarp();
    const uint32_t magicL = (uint32_t)((0x100000000ull + (uint64_t)L - 1) / (uint64_t)L);   // head / L == umulhi(head, magicL)
    int pos = MT_N;                                                           // forces a refill at the first draw
    auto draw = [&]() -> uint32_t {
        if (pos >= MT_N) { mt_refill(s, t, lane); pos = 0; }
        return t[pos++];
    };

    int head = 0, tail = 0, Nb = 0, status = 0;
    int64_t Nb_tot = 0;
    uint64_t step = 0, Z = 0;
    int32_t ne = 0, nt = 0;
    // first multiple of 50 above therm: the event rule `step > therm && step % 50 == 0` (:113-114) without a 64-bit
    // modulo on every closed iteration (a lone warp pays dearly for it)
    uint64_t next50 = (therm / 50 + 1) * 50;

    for (;;) {
        // the (up to) four draws of this iteration, loaded together when they do not straddle a refill
        const bool closed = (tail == head);                                   // :112
        uint32_t r0, r1, r2, r3 = 0u;
        const int need = closed ? 4 : 3;
        if (pos + need <= MT_N) {
            r0 = t[pos]; r1 = t[pos + 1]; r2 = t[pos + 2];
            if (closed) r3 = t[pos + 3];
            pos += need;
        } else if (closed && step >= num_steps) {
            r0 = draw(); r1 = r2 = 0u;                                        // last iteration: the kick draw only
        } else {
            r0 = draw(); r1 = draw(); r2 = draw();
            if (closed) r3 = draw();
        }
        uint32_t u_dir = r0, u_coin = r1, u_acc = r2;
        if (closed) {
            while (next50 < step) next50 += 50;                               // (open iterations pass multiples unseen)
            if (step == next50) {                                             // :113-114  step > therm && step % 50 == 0
                if (ne < max_events) {
                    if (events && lane == 0) {
                        int64_t *e = events + ((int64_t)c * max_events + ne) * 3;
                        e[0] = (int64_t)step; e[1] = (int64_t)Z; e[2] = Nb_tot;
                    }
                    if (bond_flag && dumps) {                                 // :120-124
                        __syncwarp();
                        uint8_t *d = dumps + ((int64_t)c * max_events + ne) * nb2;
                        for (int b = lane; b < nb2; b += 32) d[b] = bonds[b];
                    }
                }
                ++ne;
            }
            if (trace && nt < max_trace && lane == 0) {
                int64_t *tr = trace + ((int64_t)c * max_trace + nt) * 2;
                tr[0] = (int64_t)step; tr[1] = Nb;
            }
            ++nt;
            // (int)floor(genrand()*N) == (u32*N) >> 32: u32*N < 2^53 is exact in double    :128
            tail = (int)(((uint64_t)r0 * (uint64_t)N) >> 32);
            head = tail;
            Z += 1;                                                           // :130
            Nb_tot -= Nb;                                                     // :132
            if (step >= num_steps) {                                          // :133-134
                break;                                                        // (draws loaded ahead of a shift that never happens change nothing observable)
            }
            u_dir = r1; u_coin = r2; u_acc = r3;
        }
        const int d = (int)(u_dir >> 30);                                     // floor(genrand()*4)  :137
        const int y = (int)fastdiv((uint32_t)head, magicL), x = head - y * L;
        int nh, ib;
        if (L == 2) {
            // bond_number's branch order sends both x-directions to the x == 0 site's bond (likewise y)
            if (d == 0)      { nh = y * L + (x + 1) % L;          ib = 2 * (y * L); }
            else if (d == 1) { nh = ((y + 1) % L) * L + x;        ib = 2 * x + 1; }
            else if (d == 2) { nh = y * L + (x + L - 1) % L;      ib = 2 * (y * L); }
            else             { nh = ((y + L - 1) % L) * L + x;    ib = 2 * x + 1; }
        } else {
            // neighbour table :88-95 and bond_number :214-244 in closed form
            const bool ym = d & 1, neg = d & 2;
            const int cc = ym ? y : x;
            const int edge = neg ? 0 : L - 1;
            const int diff = (cc == edge) ? (neg ? L - 1 : 1 - L) : (neg ? -1 : 1);
            nh = head + diff * (ym ? L : 1);
            ib = 2 * (neg ? nh : head) + (int)ym;
        }
        const int nb = bonds[ib];                                             // :139
        const bool inc = u_coin < 0x80000000u;                                // genrand() < 0.5  :141
        const int delta = inc ? 1 : -1;
        unsigned long long thr;
        if (nb < THR_N) thr = inc ? thr_inc[nb] : thr_dec[nb];
        else thr = inc ? thr_of(p.K / ((double)nb + 1.0)) : thr_of((double)nb * p.inv_K);
        // genrand() < P  :151.  Every lane walks the chain redundantly and stores the same byte itself, so a lane only ever
        // reads what it has written: no warp synchronisation, and the update stays predicated (no branch).
        const bool ok = (unsigned long long)u_acc < thr;
        const bool over = ok && (nb + delta > 255);
        const bool acc = ok && !over;
        status = over ? 1 : status;
        if (acc) bonds[ib] = (uint8_t)(nb + delta);
        Nb += acc ? delta : 0;
        head = acc ? nh : head;
        ++step;                                                               // :156
    }
    __syncwarp();
    if (SMEM) {
        for (int b = lane; b < nb2; b += 32) gb[b] = bonds[b];
    }
    if (lane == 0) {
        int64_t *o = out + c * MT_STRIDE;
        o[0] = (int64_t)Z; o[1] = Nb_tot; o[2] = (int64_t)step; o[3] = Nb; o[4] = status;
        o[5] = 0; o[6] = 0; o[7] = 0;
        if (n_events) n_events[c] = ne;
        if (n_trace) n_trace[c] = nt;
    }
}

}  // namespace

cudaError_t launch_mt(int L, int64_t n_chains, int64_t n_pad, const MtChain *d_prm,
                      uint64_t num_steps, uint64_t therm, int bond_flag,
                      uint8_t *d_bonds, int64_t *d_out,
                      int32_t max_events, int32_t *d_n_events, int64_t *d_events, uint8_t *d_dumps,
                      int32_t max_trace, int32_t *d_n_trace, int64_t *d_trace,
                      uint32_t *d_mt_state, int smem_optin, cudaStream_t st)
{
    (void)n_pad; (void)d_mt_state;
    const size_t base = (size_t)2 * MT_N * 4 + (size_t)2 * THR_N * 8;
    const size_t with_bonds = base + (size_t)2 * L * L;
    const bool smem = with_bonds + 1024 <= (size_t)smem_optin;
    const size_t bytes = smem ? with_bonds : base;
    auto kern = smem ? worm_mt_kernel<true> : worm_mt_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)n_chains, 32, bytes, st>>>(L, n_chains, d_prm, num_steps, therm, bond_flag, d_bonds, d_out,
                                               max_events, d_n_events, d_events, d_dumps,
                                               max_trace, d_n_trace, d_trace);
    return cudaGetLastError();
}

}  // namespace worm
