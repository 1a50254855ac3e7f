// worm_philox.cu -- production worm kernels for sm_100a (hand-written CUDA, no tensor cores:
// the path is a chain of dependent integer operations on a few KB of state per Markov chain).
//
// Algorithm (restated on the CPU in oracle/worm_oracle.c:worm_oracle_philox_run and checked
// bit-for-bit against it): the reference loop worm_ising_2d.cpp:111-157 with
//   - a counter-based Philox4x32-10 draw per step, counter = {step, global chain id}, key = seed,
//     words {kick, direction, +/- coin, acceptance} in the reference's draw order (:128,137,141,151);
//   - integer acceptance tests  (nb+1)*u < kfix  and  u < nb*tfix  for K/(nb+1) and nb/K (:143,:148);
//   - measurement of every closed iteration (:130-132) once step >= therm.
//
// Mapping.  A chain is owned by a group of G lanes of one warp (G = 1: a thread per chain, most
// issue-efficient when there are enough chains to fill the machine; G = 4: a quad per chain, the
// quad's lanes generate the Philox words of 4 consecutive steps in parallel and all walk the
// (sequential) chain, which shortens the per-step issue cost of a chain ~2x when the chain count,
// not the issue rate, limits throughput).  A block is ONE warp; the bond occupations of its 32/G
// chains are staged in shared memory for the whole launch, word-interleaved across chains
// ([word][chain], so the lanes of a warp always hit distinct banks), and written back at the end.
// Lattices too large for shared memory run the same code on the global-memory copy (L1/L2 resident).
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

namespace {

constexpr int ALGO_TAYLOR = 0;
constexpr int ALGO_BINARY = 1;

struct Geo {
    int L, N, nb2;
    uint32_t magicL;     // ceil(2^32 / L)
    int words;           // 32-bit words of bond storage per chain in shared memory
    int ch;              // chains per block
};

// ---------------------------------------------------------------------------------------------
// Bond storage.  SMEM Taylor: uint8 per bond, 4 bonds per word, word w of chain slot q at
// smem32[w*ch + q].  SMEM binary: 1 bit per bond, 32 per word, same interleave.  GLOBAL: the
// caller's uint8 [2N] array of the chain (one byte per bond in both algorithms).
// ---------------------------------------------------------------------------------------------
template <int ALGO, bool SMEM>
struct Store {
    uint8_t *g;          // GLOBAL: chain's bonds
    uint32_t *s;         // SMEM: &smem32[q]
    int ch;

    __device__ __forceinline__ uint32_t load(int ib, uint32_t &aux) const
    {
        if constexpr (!SMEM) {
            aux = 0;
            return g[ib];
        } else if constexpr (ALGO == ALGO_TAYLOR) {
            const uint8_t *p = reinterpret_cast<const uint8_t *>(s + (ib >> 2) * ch) + (ib & 3);
            aux = 0;
            return *p;
        } else {
            aux = s[(ib >> 5) * ch];
            return (aux >> (ib & 31)) & 1u;
        }
    }
    __device__ __forceinline__ void store(int ib, uint32_t v, uint32_t aux) const
    {
        if constexpr (!SMEM) {
            g[ib] = (uint8_t)v;
        } else if constexpr (ALGO == ALGO_TAYLOR) {
            uint8_t *p = reinterpret_cast<uint8_t *>(s + (ib >> 2) * ch) + (ib & 3);
            *p = (uint8_t)v;
        } else {
            s[(ib >> 5) * ch] = aux ^ (1u << (ib & 31));     // accepted binary move always flips
        }
    }
    // byte view of bond ib (for dumps / write-back)
    __device__ __forceinline__ uint32_t peek(int ib) const
    {
        uint32_t aux;
        return load(ib, aux);
    }
};

struct ChainState {
    int head, tail, hx, hy, tx, ty;
    int Nb, npar;                       // sum of occupations, number of odd bonds
    uint64_t Z, SNb, SNb2, Sn, Sn2;
};

// One worm step.  MEAS: step >= therm.  `dump_now` is warp-uniform (step is a dump multiple).
template <int ALGO, bool SMEM, bool HIST, bool MEAS>
__device__ __forceinline__ void worm_step(const Geo &geo, const PhiloxChain &prm, const Store<ALGO, SMEM> &st,
                                          ChainState &c, const uint4 r, bool live, bool writer,
                                          int64_t *hist_row)
{
    const int L = geo.L;
    const bool closed = (c.head == c.tail);
    if constexpr (MEAS) {
        if constexpr (HIST) {
            if (writer && hist_row) {
                int dx = c.hx - c.tx; dx += (dx < 0) ? L : 0;
                int dy = c.hy - c.ty; dy += (dy < 0) ? L : 0;
                atomicAdd(reinterpret_cast<unsigned long long *>(hist_row + dy * L + dx), 1ull);
            }
        }
        if (closed) {                                           // worm_ising_2d.cpp:130-132
            c.Z += 1;
            c.SNb += (uint32_t)c.Nb;
            c.SNb2 += (uint64_t)(uint32_t)c.Nb * (uint32_t)c.Nb;
            c.Sn += (uint32_t)c.npar;
            c.Sn2 += (uint64_t)(uint32_t)c.npar * (uint32_t)c.npar;
        }
    }
    if (closed) {                                               // kick :128-129
        const int site = (int)__umulhi(r.x, (uint32_t)geo.N);
        const int ky = (int)fastdiv((uint32_t)site, geo.magicL);
        const int kx = site - ky * L;
        c.head = c.tail = site;
        c.hx = c.tx = kx;
        c.hy = c.ty = ky;
    }
    // shift move :137-155
    const uint32_t d = r.y >> 30;
    const bool ym = d & 1u, neg = d & 2u;
    const int cc = ym ? c.hy : c.hx;
    const int edge = neg ? 0 : L - 1;
    const int wrapd = neg ? L - 1 : 1 - L;
    const int normd = neg ? -1 : 1;
    const int diff = (cc == edge) ? wrapd : normd;
    const int stride = ym ? L : 1;
    const int nh = c.head + diff * stride;
    const int osite = neg ? nh : c.head;                        // bond_number :214-244 (L >= 3)
    const int ib = 2 * osite + (int)ym;
    uint32_t aux;
    const uint32_t nb = st.load(ib, aux);
    bool acc;
    int delta;
    if constexpr (ALGO == ALGO_TAYLOR) {
        const bool inc = (int32_t)r.z >= 0;                     // genrand() < 0.5  :141
        const bool a_inc = (nb < 255u) && ((uint64_t)(nb + 1u) * r.w < prm.kfix);   // u < K/(nb+1)  :143,151
        const bool a_dec = (uint64_t)r.w < (uint64_t)nb * prm.tfix;                  // u < nb/K      :148,151
        acc = inc ? a_inc : a_dec;
        delta = inc ? 1 : -1;
    } else {
        acc = nb ? true : ((uint64_t)r.w < prm.hfix);           // tanh(K) for an empty bond, 1 otherwise
        delta = nb ? -1 : 1;
    }
    if (acc) {
        const uint32_t nn = nb + delta;
        // every live lane of the group stores the (identical) value, so each lane's later loads are
        // ordered after its own store; dead lanes never write
        if (live) st.store(ib, nn, aux);
        c.Nb += delta;
        c.npar += (nn & 1u) ? 1 : -1;
        c.head = nh;
        if (ym) c.hy += diff; else c.hx += diff;
    }
}

template <int ALGO, bool SMEM>
__device__ void dump_chain(const Geo &geo, const Store<ALGO, SMEM> &st, uint8_t *dst)
{
    for (int b = 0; b < geo.nb2; ++b) dst[b] = (uint8_t)st.peek(b);
}

// ---------------------------------------------------------------------------------------------
// The kernel: one warp per block, 32/G chain slots.
// ---------------------------------------------------------------------------------------------
template <int ALGO, int G, bool SMEM, bool HIST>
__global__ void __launch_bounds__(32)
worm_philox_kernel(const PhiloxArgs a, const Geo geo)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x;
    const int sub = lane % G;            // lane within the chain's group
    const int q = lane / G;              // chain slot within the warp
    const int64_t cidx = (int64_t)blockIdx.x * geo.ch + q;
    const bool live = (q < geo.ch) && (cidx < a.n_chains);
    const bool writer = live && (sub == 0);
    const int64_t cl = live ? cidx : 0;  // dead lanes shadow chain 0 of the launch, never write

    // shared memory: [bond words][ch] then (G > 1) the per-round Philox exchange buffer [G][32/G] uint4
    uint32_t *bond_smem = smem;
    uint4 *rbuf = reinterpret_cast<uint4 *>(smem + (((size_t)geo.words * geo.ch + 3) & ~(size_t)3));

    Store<ALGO, SMEM> st;
    st.g = a.bonds + cl * (int64_t)geo.nb2;
    st.s = bond_smem + (q < geo.ch ? q : 0);
    st.ch = geo.ch;

    const PhiloxChain prm = a.prm[cl];
    ChainState c;
    c.head = a.head[cl];
    c.tail = a.tail[cl];
    c.hy = (int)fastdiv((uint32_t)c.head, geo.magicL); c.hx = c.head - c.hy * geo.L;
    c.ty = (int)fastdiv((uint32_t)c.tail, geo.magicL); c.tx = c.tail - c.ty * geo.L;
    {
        const int64_t *ac = a.acc + cl * ACC_STRIDE;
        c.Z = (uint64_t)ac[0]; c.SNb = (uint64_t)ac[1]; c.SNb2 = (uint64_t)ac[2];
        c.Sn = (uint64_t)ac[3]; c.Sn2 = (uint64_t)ac[4];
    }
    // stage the chain's bonds and recompute Nb / parity count
    {
        int Nb = 0, npar = 0;
        if constexpr (SMEM) {
            if (live) {
                for (int w = sub; w < geo.words; w += G) {
                    uint32_t word = 0;
                    if constexpr (ALGO == ALGO_TAYLOR) {
                        for (int k = 0; k < 4; ++k) {
                            const int b = 4 * w + k;
                            const uint32_t v = (b < geo.nb2) ? st.g[b] : 0u;
                            word |= v << (8 * k);
                        }
                    } else {
                        for (int k = 0; k < 32; ++k) {
                            const int b = 32 * w + k;
                            const uint32_t v = (b < geo.nb2) ? (st.g[b] & 1u) : 0u;
                            word |= v << k;
                        }
                    }
                    st.s[w * geo.ch] = word;
                }
            }
            __syncwarp();
        }
        for (int b = 0; b < geo.nb2; ++b) {          // every lane of the group needs the totals
            const uint32_t v = SMEM ? st.peek(b) : (uint32_t)st.g[b];
            Nb += (int)v; npar += (int)(v & 1u);
        }
        c.Nb = Nb; c.npar = npar;
    }
    int64_t *hist_row = nullptr;
    if constexpr (HIST) {
        if (a.hist && a.group_index) hist_row = a.hist + (int64_t)a.group_index[cl] * geo.N;
    }
    int32_t nd = (a.n_dumps && live) ? a.n_dumps[cl] : 0;

    const uint64_t chain_id = a.chain_id0 + (uint64_t)cl;
    const uint32_t k0 = (uint32_t)prm.seed, k1 = (uint32_t)(prm.seed >> 32);
    const uint32_t c2 = (uint32_t)chain_id, c3 = (uint32_t)(chain_id >> 32);

    // Segment the step range so that (a) MEAS is a compile-time constant inside a segment and
    // (b) a dump multiple can only be the FIRST step of a segment.
    uint64_t s = a.step_begin;
    while (s < a.step_end) {
        const bool meas = (s >= a.therm);
        uint64_t seg_end = a.step_end;
        if (!meas && a.therm < seg_end) seg_end = a.therm;
        bool dump_first = false;
        if (meas && a.dump_every) {
            const uint64_t rem = s % a.dump_every;
            dump_first = (rem == 0);
            const uint64_t nxt = s + (a.dump_every - rem);       // next multiple after s
            if (nxt < seg_end) seg_end = nxt;
        }
        if (dump_first && live && c.head == c.tail) {            // worm_ising_2d.cpp:113-124
            if (writer) {
                if (nd < a.max_dumps) {
                    if (a.events) {
                        int64_t *e = a.events + ((int64_t)cl * a.max_dumps + nd) * 3;
                        e[0] = (int64_t)s; e[1] = (int64_t)c.Z; e[2] = (int64_t)c.SNb;
                    }
                    if (a.dumps) dump_chain<ALGO, SMEM>(geo, st, a.dumps + ((int64_t)cl * a.max_dumps + nd) * geo.nb2);
                }
            }
            ++nd;
        }
        if constexpr (G == 1) {
            if (meas) {
                for (; s < seg_end; ++s) {
                    const uint4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), c2, c3, k0, k1);
                    worm_step<ALGO, SMEM, HIST, true>(geo, prm, st, c, r, live, writer, hist_row);
                }
            } else {
                for (; s < seg_end; ++s) {
                    const uint4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), c2, c3, k0, k1);
                    worm_step<ALGO, SMEM, HIST, false>(geo, prm, st, c, r, live, writer, hist_row);
                }
            }
        } else {
            // rounds of G steps: lane `sub` draws the words of step s+sub, the group exchanges them
            // through shared memory and every lane of the group walks the chain through all of them
            uint4 *my = rbuf + q;                     // rbuf[j * (32/G) + q]
            constexpr int RS = 32 / G;
            while (s < seg_end) {
                const uint64_t sj = s + (uint64_t)sub;
                const uint4 rr = philox4x32_10((uint32_t)sj, (uint32_t)(sj >> 32), c2, c3, k0, k1);
                __syncwarp();
                my[sub * RS] = rr;
                __syncwarp();
                const uint64_t left = seg_end - s;
                const int nj = left < (uint64_t)G ? (int)left : G;
                if (meas) {
                    if (nj == G) {
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            worm_step<ALGO, SMEM, HIST, true>(geo, prm, st, c, my[j * RS], live, writer, hist_row);
                    } else {
                        for (int j = 0; j < nj; ++j)
                            worm_step<ALGO, SMEM, HIST, true>(geo, prm, st, c, my[j * RS], live, writer, hist_row);
                    }
                } else {
                    if (nj == G) {
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            worm_step<ALGO, SMEM, HIST, false>(geo, prm, st, c, my[j * RS], live, writer, hist_row);
                    } else {
                        for (int j = 0; j < nj; ++j)
                            worm_step<ALGO, SMEM, HIST, false>(geo, prm, st, c, my[j * RS], live, writer, hist_row);
                    }
                }
                s += (uint64_t)nj;
            }
        }
    }

    // write back
    if constexpr (SMEM) {
        __syncwarp();
        if (live) {
            for (int b = sub; b < geo.nb2; b += G) st.g[b] = (uint8_t)st.peek(b);
        }
    }
    if (writer) {
        a.head[cl] = c.head;
        a.tail[cl] = c.tail;
        int64_t *ac = a.acc + cl * ACC_STRIDE;
        const uint64_t mb = a.step_begin > a.therm ? a.step_begin : a.therm;
        const int64_t iters = (a.step_end > mb) ? (int64_t)(a.step_end - mb) : 0;
        if (a.group_acc && a.group_index) {
            unsigned long long *ga = reinterpret_cast<unsigned long long *>(a.group_acc + (int64_t)a.group_index[cl] * ACC_STRIDE);
            atomicAdd(ga + 0, (unsigned long long)(c.Z - (uint64_t)ac[0]));
            atomicAdd(ga + 1, (unsigned long long)(c.SNb - (uint64_t)ac[1]));
            atomicAdd(ga + 2, (unsigned long long)(c.SNb2 - (uint64_t)ac[2]));
            atomicAdd(ga + 3, (unsigned long long)(c.Sn - (uint64_t)ac[3]));
            atomicAdd(ga + 4, (unsigned long long)(c.Sn2 - (uint64_t)ac[4]));
            atomicAdd(ga + 5, (unsigned long long)iters);
        }
        ac[0] = (int64_t)c.Z; ac[1] = (int64_t)c.SNb; ac[2] = (int64_t)c.SNb2;
        ac[3] = (int64_t)c.Sn; ac[4] = (int64_t)c.Sn2;
        ac[5] += iters;
        if (a.n_dumps) a.n_dumps[cl] = nd;
    }
}

template <int ALGO, int G, bool SMEM, bool HIST>
cudaError_t launch_one(const PhiloxArgs &a, const Geo &geo, size_t smem_bytes, cudaStream_t st)
{
    auto kern = worm_philox_kernel<ALGO, G, SMEM, HIST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (a.n_chains + geo.ch - 1) / geo.ch;
    kern<<<(unsigned)blocks, 32, smem_bytes, st>>>(a, geo);
    return cudaGetLastError();
}

template <int ALGO, int G>
cudaError_t launch_g(const PhiloxArgs &a, const Geo &geo, bool smem, bool hist, size_t smem_bytes, cudaStream_t st)
{
    if (smem) {
        if (hist) return launch_one<ALGO, G, true, true>(a, geo, smem_bytes, st);
        return launch_one<ALGO, G, true, false>(a, geo, smem_bytes, st);
    }
    if (hist) return launch_one<ALGO, G, false, true>(a, geo, smem_bytes, st);
    return launch_one<ALGO, G, false, false>(a, geo, smem_bytes, st);
}

}  // namespace

cudaError_t launch_philox(const PhiloxArgs &a, int lanes_per_chain, int smem_optin, int sm_count,
                          cudaStream_t st, const char **why)
{
    *why = "";
    Geo geo;
    geo.L = a.L;
    geo.N = a.L * a.L;
    geo.nb2 = 2 * geo.N;
    geo.magicL = (uint32_t)((0x100000000ull + (uint64_t)a.L - 1) / (uint64_t)a.L);
    geo.words = (a.algo == ALGO_TAYLOR) ? (geo.nb2 + 3) / 4 : (geo.nb2 + 31) / 32;

    int G = lanes_per_chain;
    const bool force_global = (G == -1);
    if (force_global) G = 1;
    if (G == 0) {
        // quad-per-chain while the chain count cannot fill the issue slots with a thread per chain
        const int64_t fill = (int64_t)sm_count * 4 /*SMSPs*/ * 8 /*chains per quad-warp*/ * 2;
        G = (a.n_chains <= fill) ? 4 : 1;
    }
    if (G != 1 && G != 4) { *why = "lanes_per_chain must be 0, 1 or 4"; return cudaErrorInvalidValue; }

    int slots = 32 / G;
    const size_t per_chain = (size_t)geo.words * 4;
    size_t rbuf_bytes = (G > 1) ? 32 * sizeof(uint4) + 16 : 0;
    const size_t budget = (size_t)smem_optin > rbuf_bytes + 1024 ? (size_t)smem_optin - rbuf_bytes - 1024 : 0;
    int ch = (int)(budget / per_chain);
    if (ch > slots) ch = slots;
    const bool smem = ch >= 1 && !force_global;
    if (!smem) { G = 1; ch = 32; rbuf_bytes = 0; }   // global-memory bonds: one thread per chain
    geo.ch = ch;
    const size_t smem_bytes = smem ? (((size_t)geo.words * ch * 4 + 15) & ~(size_t)15) + rbuf_bytes
                                   : rbuf_bytes;
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    if (a.algo == ALGO_TAYLOR) {
        if (G == 1) return launch_g<ALGO_TAYLOR, 1>(a, geo, smem, hist, smem_bytes, st);
        return launch_g<ALGO_TAYLOR, 4>(a, geo, smem, hist, smem_bytes, st);
    }
    if (G == 1) return launch_g<ALGO_BINARY, 1>(a, geo, smem, hist, smem_bytes, st);
    return launch_g<ALGO_BINARY, 4>(a, geo, smem, hist, smem_bytes, st);
}

}  // namespace worm
