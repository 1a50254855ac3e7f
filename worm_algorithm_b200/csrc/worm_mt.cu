// worm_mt.cu -- deterministic mode: the reference's Markov chain on the reference's random stream.
//
// One thread per chain.  Each thread walks the loop of worm_ising_2d.cpp:111-157 literally --
// the same draw order (kick :128, direction :137, +/- coin :141, acceptance :151), the same
// double-precision acceptance ratio K/(nb+1.0) or nb*inv_K (:143,:148; IEEE division and
// multiplication round identically on the device), the same stop rule (:133-134) -- on its own
// MT19937 state, seeded as mt19937ar.c:68-81.  The 624-word state lives in global memory,
// word-major ([624][n_pad]) so a warp's refill is coalesced; bonds are uint8 in d_bonds.
//
// This kernel exists for bit-exact parity, not speed: chains stop at chain-specific steps and
// consume draws conditionally, so lanes diverge.  The production path is worm_philox.cu.
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

struct MtGen {
    uint32_t *s;      // this chain's column of the [624][n_pad] state
    int64_t stride;   // n_pad
    int pos;

    __device__ __forceinline__ uint32_t &at(int i) { return s[(int64_t)i * stride]; }

    __device__ void seed(uint32_t v)
    {
        at(0) = v;
        uint32_t p = v;
        for (int i = 1; i < 624; ++i) {
            p = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
            at(i) = p;
        }
        pos = 624;
    }

    __device__ void refill()
    {
        // in-place twist; index arithmetic mod 624 reproduces the three-loop form of
        // mt19937ar.c:121-134 (s[i+397] for i >= 227 is an already-updated word, as there)
        uint32_t cur = at(0);
        const uint32_t first_old = cur;
        for (int i = 0; i < 624; ++i) {
            const uint32_t nxt = (i + 1 < 624) ? at(i + 1) : at(0);
            const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
            const int j = (i + 397 < 624) ? i + 397 : i + 397 - 624;
            uint32_t v = at(j) ^ (y >> 1);
            if (y & 1u) v ^= 0x9908b0dfu;
            at(i) = v;
            cur = nxt;
        }
        (void)first_old;
        pos = 0;
    }

    __device__ __forceinline__ uint32_t next()
    {
        if (pos >= 624) refill();
        uint32_t y = at(pos++);
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

__global__ void __launch_bounds__(64)
worm_mt_kernel(int L, int64_t n_chains, int64_t n_pad, const MtChain *__restrict__ prm,
               uint64_t num_steps, uint64_t therm, int bond_flag,
               uint8_t *__restrict__ bonds_all, int64_t *__restrict__ out,
               int32_t max_events, int32_t *__restrict__ n_events, int64_t *__restrict__ events,
               uint8_t *__restrict__ dumps,
               int32_t max_trace, int32_t *__restrict__ n_trace, int64_t *__restrict__ trace,
               uint32_t *__restrict__ mt_state)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    const int N = L * L;
    const int nb2 = 2 * N;
    const MtChain p = prm[c];
    uint8_t *bonds = bonds_all + c * (int64_t)nb2;
    for (int b = 0; b < nb2; ++b) bonds[b] = 0;                       // worm_ising_2d.cpp:97-101

    MtGen g;
    g.s = mt_state + c;
    g.stride = n_pad;
    g.seed(p.seed);                                                    // :47

    int head = 0, tail = 0, Nb = 0, status = 0;
    int64_t Nb_tot = 0;
    uint64_t step = 0, Z = 0;
    int32_t ne = 0, nt = 0;
    const double two_m32 = 1.0 / 4294967296.0;                         // mt19937ar.c:187

    for (;;) {
        if (tail == head) {                                            // :112
            if (step > therm && step % 50 == 0) {                      // :113-114
                if (ne < max_events) {
                    if (events) {
                        int64_t *e = events + ((int64_t)c * max_events + ne) * 3;
                        e[0] = (int64_t)step; e[1] = (int64_t)Z; e[2] = Nb_tot;
                    }
                    if (bond_flag && dumps) {                          // :120-124
                        uint8_t *d = dumps + ((int64_t)c * max_events + ne) * nb2;
                        for (int b = 0; b < nb2; ++b) d[b] = bonds[b];
                    }
                }
                ++ne;
            }
            if (trace && nt < max_trace) {
                int64_t *t = trace + ((int64_t)c * max_trace + nt) * 2;
                t[0] = (int64_t)step; t[1] = Nb;
            }
            ++nt;
            // (int)floor(genrand()*N) == (u32*N) >> 32: u32*N < 2^53 is exact in double    :128
            tail = (int)(((uint64_t)g.next() * (uint64_t)N) >> 32);
            head = tail;
            Z += 1;                                                    // :130
            Nb_tot -= Nb;                                              // :132
            if (step >= num_steps) break;                              // :133-134
        }
        const int d = (int)(g.next() >> 30);                           // floor(genrand()*4)  :137
        const int x = head % L, y = head / L;
        int nh, ib;
        // neighbour table :88-95 and bond_number :214-244 in closed form; the L == 2 branch
        // order of bond_number sends both x-directions to the x == 0 site's bond (likewise y)
        if (d == 0)      { nh = y * L + (x + 1) % L;          ib = 2 * ((L == 2) ? y * L : head); }
        else if (d == 1) { nh = ((y + 1) % L) * L + x;        ib = 2 * ((L == 2) ? x : head) + 1; }
        else if (d == 2) { nh = y * L + (x + L - 1) % L;      ib = 2 * ((L == 2) ? y * L : nh); }
        else             { nh = ((y + L - 1) % L) * L + x;    ib = 2 * ((L == 2) ? x : nh) + 1; }
        const int nb = bonds[ib];                                      // :139
        int delta;
        double P;
        if ((double)g.next() * two_m32 < 0.5) { delta = 1;  P = p.K / ((double)nb + 1.0); }     // :141-144
        else                                  { delta = -1; P = (double)nb * p.inv_K; }          // :146-149
        if ((double)g.next() * two_m32 < P) {                          // :151
            if (nb + delta > 255) { status = 1; }
            else {
                bonds[ib] = (uint8_t)(nb + delta);
                Nb += delta;
                head = nh;
            }
        }
        ++step;                                                        // :156
    }
    int64_t *o = out + c * MT_STRIDE;
    o[0] = (int64_t)Z; o[1] = Nb_tot; o[2] = (int64_t)step; o[3] = Nb; o[4] = status;
    o[5] = 0; o[6] = 0; o[7] = 0;
    if (n_events) n_events[c] = ne;
    if (n_trace) n_trace[c] = nt;
}

cudaError_t launch_mt(int L, int64_t n_chains, int64_t n_pad, const MtChain *d_prm,
                      uint64_t num_steps, uint64_t therm, int bond_flag,
                      uint8_t *d_bonds, int64_t *d_out,
                      int32_t max_events, int32_t *d_n_events, int64_t *d_events, uint8_t *d_dumps,
                      int32_t max_trace, int32_t *d_n_trace, int64_t *d_trace,
                      uint32_t *d_mt_state, cudaStream_t st)
{
    const int threads = 32;
    const int64_t blocks = (n_chains + threads - 1) / threads;
    worm_mt_kernel<<<(unsigned)blocks, threads, 0, st>>>(L, n_chains, n_pad, d_prm, num_steps, therm, bond_flag,
                                                         d_bonds, d_out, max_events, d_n_events, d_events, d_dumps,
                                                         max_trace, d_n_trace, d_trace, d_mt_state);
    return cudaGetLastError();
}

}  // namespace worm
