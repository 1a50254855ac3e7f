// worm_wide.cu -- production worm kernel, "throughput mode" (sm_100a): one thread per Markov chain.
//
// For batches with (many) more chains than the GPU has issue slots.  Bonds live in shared memory,
// word-interleaved across the 32 chains of the one-warp block ([word][chain]: the lanes of a warp
// always hit distinct banks), uint8 per bond for the Taylor chain and 1 bit per bond for the binary
// chain; lattices too large for that run on the caller's global-memory array (L1/L2 resident).
// Any L >= 3.  The algorithm and the random stream are those of worm_lat.cu and of the CPU
// restatement oracle/worm_oracle.c:worm_oracle_philox_run (bit-for-bit).
#include "worm_run.cuh"

namespace worm {

namespace {

struct Geo {
    int L, N, nb2;
    uint32_t magicL;     // ceil(2^32 / L)
    int words;           // 32-bit words of bond storage per chain in shared memory
    int ch;              // chains per block
};

// Bond storage.  SMEM Taylor: uint8 per bond, 4 bonds per word, word w of chain slot q at
// smem32[w*ch + q].  SMEM binary: 1 bit per bond, 32 per word, same interleave.  GLOBAL: the
// caller's uint8 [2N] array of the chain (one byte per bond in both algorithms).
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

template <int ALGO, bool SMEM, bool HIST, bool MEAS>
__device__ __forceinline__ void worm_step(const Geo &geo, const PhiloxChain &prm, const Store<ALGO, SMEM> &st,
                                          ChainState &c, const uint32_t ra, const uint32_t rb, bool live,
                                          int64_t *hist_row)
{
    const int L = geo.L;
    const bool closed = (c.head == c.tail);
    if constexpr (MEAS) {
        if constexpr (HIST) {
            if (live && hist_row) {
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
        const int site = (int)__umulhi(rb << 3, (uint32_t)geo.N);
        const int ky = (int)fastdiv((uint32_t)site, geo.magicL);
        const int kx = site - ky * L;
        c.head = c.tail = site;
        c.hx = c.tx = kx;
        c.hy = c.ty = ky;
    }
    // shift move :137-155
    const uint32_t d = rb >> 30;
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
        const bool inc = ((rb >> 29) & 1u) == 0;                // genrand() < 0.5  :141
        const bool a_inc = (nb < 255u) && ((uint64_t)(nb + 1u) * ra < prm.kfix);   // u < K/(nb+1)  :143,151
        const bool a_dec = (uint64_t)ra < (uint64_t)nb * prm.tfix;                  // u < nb/K      :148,151
        acc = inc ? a_inc : a_dec;
        delta = inc ? 1 : -1;
    } else {
        acc = nb ? true : ((uint64_t)ra < prm.hfix);            // tanh(K) for an empty bond, 1 otherwise
        delta = nb ? -1 : 1;
    }
    if (acc) {
        const uint32_t nn = nb + delta;
        if (live) st.store(ib, nn, aux);
        c.Nb += delta;
        c.npar += (nn & 1u) ? 1 : -1;
        c.head = nh;
        if (ym) c.hy += diff; else c.hx += diff;
    }
}

template <int ALGO, bool SMEM, bool HIST>
__global__ void __launch_bounds__(32)
worm_wide_kernel(const PhiloxArgs a, const Geo geo)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int q = threadIdx.x;           // chain slot within the warp
    const int64_t cidx = (int64_t)blockIdx.x * geo.ch + q;
    const bool live = (q < geo.ch) && (cidx < a.n_chains);
    const int64_t cl = live ? cidx : 0;  // dead lanes shadow chain 0 of the launch, never write

    Store<ALGO, SMEM> st;
    st.g = a.bonds + cl * (int64_t)geo.nb2;
    st.s = smem + (q < geo.ch ? q : 0);
    st.ch = geo.ch;

    const PhiloxChain prm = a.prm[cl];
    ChainState c;
    c.head = a.head[cl];
    c.tail = a.tail[cl];
    c.hy = (int)fastdiv((uint32_t)c.head, geo.magicL); c.hx = c.head - c.hy * geo.L;
    c.ty = (int)fastdiv((uint32_t)c.tail, geo.magicL); c.tx = c.tail - c.ty * geo.L;
    const int64_t *ac0 = a.acc + cl * ACC_STRIDE;
    c.Z = (uint64_t)ac0[0]; c.SNb = (uint64_t)ac0[1]; c.SNb2 = (uint64_t)ac0[2];
    c.Sn = (uint64_t)ac0[3]; c.Sn2 = (uint64_t)ac0[4];
    // stage the chain's bonds and recompute Nb / parity count
    {
        int Nb = 0, npar = 0;
        if constexpr (SMEM) {
            if (live) {
                for (int w = 0; w < geo.words; ++w) {
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
        for (int b = 0; b < geo.nb2; ++b) {
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

    Segments segs(a);
    uint64_t s = a.step_begin;
    while (s < a.step_end) {
        bool meas, dump_first;
        const uint64_t seg_end = segs.next(a, s, meas, dump_first);
        if (dump_first && live && c.head == c.tail) {            // worm_ising_2d.cpp:113-124
            if (nd < a.max_dumps) {
                if (a.events) {
                    int64_t *e = a.events + ((int64_t)cl * a.max_dumps + nd) * 3;
                    e[0] = (int64_t)s; e[1] = (int64_t)c.Z; e[2] = (int64_t)c.SNb;
                }
                if (a.dumps) {
                    uint8_t *dst = a.dumps + ((int64_t)cl * a.max_dumps + nd) * geo.nb2;
                    for (int b = 0; b < geo.nb2; ++b) dst[b] = (uint8_t)st.peek(b);
                }
            }
            ++nd;
        }
        // one Philox block per pair of steps; a segment may begin or end inside a pair
        while (s < seg_end) {
            const uint64_t pair = s >> 1;
            const uint4 r = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), c2, c3, k0, k1);
            uint32_t ra, rb;
            if ((s & 1ull) == 0) {
                step_words(r, false, ra, rb);
                if (meas) worm_step<ALGO, SMEM, HIST, true>(geo, prm, st, c, ra, rb, live, hist_row);
                else worm_step<ALGO, SMEM, HIST, false>(geo, prm, st, c, ra, rb, live, hist_row);
                if (++s >= seg_end) break;
            }
            step_words(r, true, ra, rb);
            if (meas) worm_step<ALGO, SMEM, HIST, true>(geo, prm, st, c, ra, rb, live, hist_row);
            else worm_step<ALGO, SMEM, HIST, false>(geo, prm, st, c, ra, rb, live, hist_row);
            ++s;
        }
    }

    // write back
    if constexpr (SMEM) {
        __syncwarp();
        if (live) {
            for (int b = 0; b < geo.nb2; ++b) st.g[b] = (uint8_t)st.peek(b);
        }
    }
    if (live) {
        a.head[cl] = c.head;
        a.tail[cl] = c.tail;
        int64_t *ac = a.acc + cl * ACC_STRIDE;
        const int64_t iters = measured_iters(a);
        add_group_acc(a, cl, ac, c.Z, c.SNb, c.SNb2, c.Sn, c.Sn2, iters);
        ac[0] = (int64_t)c.Z; ac[1] = (int64_t)c.SNb; ac[2] = (int64_t)c.SNb2;
        ac[3] = (int64_t)c.Sn; ac[4] = (int64_t)c.Sn2;
        ac[5] += iters;
        if (a.n_dumps) a.n_dumps[cl] = nd;
    }
}

template <int ALGO, bool SMEM, bool HIST>
cudaError_t launch_wide(const PhiloxArgs &a, const Geo &geo, size_t smem_bytes, cudaStream_t st)
{
    auto kern = worm_wide_kernel<ALGO, SMEM, HIST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (a.n_chains + geo.ch - 1) / geo.ch;
    kern<<<(unsigned)blocks, 32, smem_bytes, st>>>(a, geo);
    return cudaGetLastError();
}

template <int ALGO>
cudaError_t launch_wide_a(const PhiloxArgs &a, const Geo &geo, bool smem, bool hist, size_t smem_bytes, cudaStream_t st)
{
    if (smem) {
        if (hist) return launch_wide<ALGO, true, true>(a, geo, smem_bytes, st);
        return launch_wide<ALGO, true, false>(a, geo, smem_bytes, st);
    }
    if (hist) return launch_wide<ALGO, false, true>(a, geo, smem_bytes, st);
    return launch_wide<ALGO, false, false>(a, geo, smem_bytes, st);
}

}  // namespace

cudaError_t launch_philox_wide(const PhiloxArgs &a, bool force_global, int smem_optin, cudaStream_t st)
{
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    Geo geo;
    geo.L = a.L;
    geo.N = a.L * a.L;
    geo.nb2 = 2 * geo.N;
    geo.magicL = (uint32_t)((0x100000000ull + (uint64_t)a.L - 1) / (uint64_t)a.L);
    geo.words = (a.algo == ALGO_TAYLOR) ? (geo.nb2 + 3) / 4 : (geo.nb2 + 31) / 32;
    const size_t per_chain = (size_t)geo.words * 4;
    const size_t budget = (size_t)smem_optin > 1024 ? (size_t)smem_optin - 1024 : 0;
    int ch = (int)(budget / per_chain);
    if (ch > 32) ch = 32;
    const bool smem = ch >= 1 && !force_global;
    if (!smem) ch = 32;
    geo.ch = ch;
    const size_t smem_bytes = smem ? (((size_t)geo.words * ch * 4 + 15) & ~(size_t)15) : 0;
    if (a.algo == ALGO_TAYLOR) return launch_wide_a<ALGO_TAYLOR>(a, geo, smem, hist, smem_bytes, st);
    return launch_wide_a<ALGO_BINARY>(a, geo, smem, hist, smem_bytes, st);
}

}  // namespace worm
