// worm_philox.cu -- production worm kernels for sm_100a (hand-written CUDA, no tensor cores:
// the path is a chain of dependent integer operations on a few KB of state per Markov chain).
//
// Algorithm (restated on the CPU in oracle/worm_oracle.c:worm_oracle_philox_run and checked
// bit-for-bit against it): the reference loop worm_ising_2d.cpp:111-157 with
//   - a counter-based Philox4x32-10 stream, one call per TWO steps (counter = {step >> 1, chain id},
//     key = seed); a step's two words are a = acceptance draw (:151) and b = direction (:137) |
//     +/- coin (:141) | kick-site bits (:128);
//   - integer acceptance tests  (nb+1)*a < kfix  and  a < nb*tfix  for K/(nb+1) and nb/K (:143,:148);
//   - measurement of every closed iteration (:130-132) once step >= therm.
//
// Two kernels, chosen by the number of chains (launch_philox):
//
//   worm_lat_kernel   "latency mode", for sweeps with fewer chains than the GPU has issue slots
//       (BASELINE configs[1]: 4096 chains on 592 SM sub-partitions).  A chain is a sequence of
//       dependent steps, so throughput = chains / (cycles per step): the kernel is built around the
//       per-step critical path  head -> bond address -> LDS -> accept -> head.  A chain is owned
//       by a group of G lanes of a one-warp block.  Each lane of the group draws the Philox words
//       of two steps of the NEXT round and turns them, off the critical path, into ready-to-use
//       step records (pre-decoded head increment / wrap mask / bond address bits / kick site /
//       occupation threshold); records go through shared memory to all lanes of the group, which
//       walk the chain redundantly (SIMD lanes are free here; warp issue slots are not).  Power-of-
//       two L only: the head is kept as an un-normalised doubled site index so that periodic wrap
//       is two LOP3s, and the next step's kick is folded into the accept select.
//
//   worm_wide_kernel  "throughput mode": one thread per chain, bonds in shared memory
//       (word-interleaved across the 32 chains of the warp: conflict-free) or, for lattices too
//       large for that, in global memory (L1/L2 resident).  Any L >= 3.
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

namespace {

constexpr int ALGO_TAYLOR = 0;
constexpr int ALGO_BINARY = 1;

// the two words of step s out of the Philox block of its pair
__device__ __forceinline__ void step_words(const uint4 r, bool odd, uint32_t &a, uint32_t &b)
{
    a = odd ? r.z : r.x;
    b = odd ? r.w : r.y;
}

// =============================================================================================
//                                   throughput mode (thread per chain)
// =============================================================================================
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

// what this launch added to the chain's record also goes to its group's record
__device__ __forceinline__ void add_group_acc(const PhiloxArgs &a, int64_t cl, const int64_t *old_acc,
                                              uint64_t Z, uint64_t SNb, uint64_t SNb2, uint64_t Sn, uint64_t Sn2,
                                              int64_t iters)
{
    if (a.group_acc && a.group_index) {
        unsigned long long *ga =
            reinterpret_cast<unsigned long long *>(a.group_acc + (int64_t)a.group_index[cl] * ACC_STRIDE);
        atomicAdd(ga + 0, (unsigned long long)(Z - (uint64_t)old_acc[0]));
        atomicAdd(ga + 1, (unsigned long long)(SNb - (uint64_t)old_acc[1]));
        atomicAdd(ga + 2, (unsigned long long)(SNb2 - (uint64_t)old_acc[2]));
        atomicAdd(ga + 3, (unsigned long long)(Sn - (uint64_t)old_acc[3]));
        atomicAdd(ga + 4, (unsigned long long)(Sn2 - (uint64_t)old_acc[4]));
        atomicAdd(ga + 5, (unsigned long long)iters);
    }
}

__device__ __forceinline__ int64_t measured_iters(const PhiloxArgs &a)
{
    const uint64_t mb = a.step_begin > a.therm ? a.step_begin : a.therm;
    return (a.step_end > mb) ? (int64_t)(a.step_end - mb) : 0;
}

// Segment [s, seg_end) of the step range: MEAS is constant inside it and a dump multiple can only
// be its FIRST step.  Returns seg_end, sets meas / dump_first.
__device__ __forceinline__ uint64_t next_segment(const PhiloxArgs &a, uint64_t s, bool &meas, bool &dump_first)
{
    meas = (s >= a.therm);
    uint64_t seg_end = a.step_end;
    if (!meas && a.therm < seg_end) seg_end = a.therm;
    dump_first = false;
    if (meas && a.dump_every) {
        const uint64_t rem = s % a.dump_every;
        dump_first = (rem == 0);
        const uint64_t nxt = s + (a.dump_every - rem);       // next multiple after s
        if (nxt < seg_end) seg_end = nxt;
    }
    return seg_end;
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

    uint64_t s = a.step_begin;
    while (s < a.step_end) {
        bool meas, dump_first;
        const uint64_t seg_end = next_segment(a, s, meas, dump_first);
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

// =============================================================================================
//                                   latency mode (G lanes per chain)
// =============================================================================================
struct LatGeo {
    int L, N, nb2;       // nb2 = 2N bytes of bond storage per chain (uint8 per bond)
    int M;               // 2N - 1: mask of a doubled site index
    int XM;              // 2L - 2: its x field
    int ch;              // chains per warp = 32 / G
    int rec_off;         // byte offsets inside dynamic shared memory
    int hist_off;
};

struct LatState {
    int head2, tail2;    // doubled site indices; head2 carries garbage above bit log2(2N)
    bool closed;
    int Nb, npar;
    uint64_t Z, SNb, SNb2, Sn, Sn2;
};

// A step record: everything about step s that does not depend on the chain's state.
//   ra = {add, keep, addneg, baseym}:  nh2  = (head2 & keep) | ((head2 + add) & ~keep)      new head
//                                      own2 = (head2 & keep) | ((head2 + addneg) & ~keep)   site owning the bond
//                                      bond byte = ((own2 & M) | baseym)                    (baseym = chain offset | y bit)
//   rb = {kick2, thr, delta, -}:       accept iff (nb < thr) != (delta < 0);  Taylor: nb += delta
template <int ALGO, bool LOWT>
__device__ __forceinline__ void make_record(const LatGeo &g, const PhiloxChain &prm, uint32_t a, uint32_t b,
                                            int chain_off, uint4 &ra, uint4 &rb)
{
    const uint32_t d = b >> 30;
    const bool ym = d & 1u, neg = d & 2u;
    const int stepsz = ym ? 2 * g.L : 2;
    const int add = neg ? -stepsz : stepsz;
    ra.x = (uint32_t)add;
    ra.y = ym ? 0u : ~(uint32_t)g.XM;
    ra.z = neg ? (uint32_t)add : 0u;
    ra.w = (uint32_t)chain_off | (uint32_t)ym;
    rb.x = 2u * __umulhi(b << 3, (uint32_t)g.N);                 // kick site :128
    if constexpr (ALGO == ALGO_TAYLOR) {
        const bool dec = (b >> 29) & 1u;                         // :141
        uint32_t qi, qd;
        if constexpr (!LOWT) {
            // T >= 1: kfix - 1 < 2^32 and tfix >= 2^32 > a
            // (nb+1)*a < kfix  <=>  nb < floor((kfix-1)/a); a == 0 accepts up to the storage cap.
            // Branch-free (the record builder shares a basic block with the walk it overlaps).
            const uint32_t num = (uint32_t)(prm.kfix - 1);
            const uint32_t q = num / (a | (uint32_t)(a == 0u));
            qi = (a == 0u) ? 255u : min(q, 255u);
            qd = 1u;                                             // a < nb*tfix      <=>  nb >= 1
        } else {
            const unsigned long long num = prm.kfix - 1;
            const unsigned long long q1 = a ? num / (unsigned long long)a : 255ull;
            const unsigned long long q2 = (unsigned long long)a / (unsigned long long)prm.tfix + 1ull;   // nb > floor(a/tfix)
            qi = (uint32_t)(q1 < 255ull ? q1 : 255ull);
            qd = (uint32_t)(q2 < 256ull ? q2 : 256ull);
        }
        rb.y = dec ? qd : qi;
        rb.z = dec ? 0xffffffffu : 1u;
    } else {
        rb.y = ((uint64_t)a < prm.hfix) ? 0u : 1u;               // accept iff nb >= thr (occupied: always)
        rb.z = 0xffffffffu;
    }
    rb.w = 0u;
}

// One worm step on pre-decoded records.  KICK: apply this step's kick here (otherwise the previous
// step already folded it into head2).  FOLD: fold the NEXT step's kick (kick2_next) into the head
// select, so that the head -> head dependency of consecutive steps is one select after the accept.
template <int ALGO, bool HIST, bool MEAS, bool KICK, bool FOLD>
__device__ __forceinline__ void lat_step(const LatGeo &g, uint8_t *bonds, uint32_t *hist, bool hist_writer,
                                         LatState &c, const uint4 ra, const uint4 rb, const int kick2_next)
{
    const int add = (int)ra.x, keep = (int)ra.y, addneg = (int)ra.z, baseym = (int)ra.w;
    const int kick2 = (int)rb.x, delta = (int)rb.z;
    const uint32_t thr = rb.y;
    const bool closed = c.closed;
    if constexpr (MEAS) {
        if constexpr (HIST) {
            // G(r) bin of (head - tail) before the kick; a closed worm is bin 0
            const int t1 = c.head2 - c.tail2;
            const int t2 = (c.head2 | g.XM | 1) - (c.tail2 & ~(g.XM | 1));     // y fields, no borrow from x
            const int bin = closed ? 0 : (((t1 & g.XM) | (t2 & g.M & ~(g.XM | 1))) >> 1);
            if (hist_writer) atomicAdd(hist + bin, 1u);
        }
        if (closed) {                                           // worm_ising_2d.cpp:130-132
            c.Z += 1;
            c.SNb += (uint32_t)c.Nb;
            c.SNb2 += (uint64_t)(uint32_t)c.Nb * (uint32_t)c.Nb;
            c.Sn += (uint32_t)c.npar;
            c.Sn2 += (uint64_t)(uint32_t)c.npar * (uint32_t)c.npar;
        }
    }
    if constexpr (KICK) c.head2 = closed ? kick2 : c.head2;     // kick :128-129
    const int tail2 = closed ? kick2 : c.tail2;
    c.tail2 = tail2;
    const int head2 = c.head2;
    const int t = head2 + add;
    const int nh2 = (head2 & keep) | (t & ~keep);               // neighbour table :88-95
    const int u = head2 + addneg;
    const int own = (head2 & keep) | (u & ~keep);               // bond_number :214-244
    const int off = (own & g.M) | baseym;
    const uint32_t nb = bonds[off];                             // :139
    const bool pc = ((nh2 ^ tail2) & g.M) == 0;                 // closed after an accepted move?
    int A = nh2, B = head2;
    if constexpr (FOLD) {
        A = pc ? kick2_next : nh2;
        B = closed ? kick2_next : head2;
    }
    // accept test (:151) and head select in one opaque block: the compiler must not re-associate
    // "closed ? kick : ..." across the select (it would put two more operations on the
    // head -> head critical path)
    int head_new;
    uint32_t acc_i;
    asm("{\n\t.reg .pred pa, pd;\n\t"
        "setp.lt.u32 pa, %2, %3;\n\t"
        "setp.lt.s32 pd, %4, 0;\n\t"
        "xor.pred pa, pa, pd;\n\t"
        "selp.b32 %0, %5, %6, pa;\n\t"
        "selp.u32 %1, 1, 0, pa;\n\t}"
        : "=r"(head_new), "=r"(acc_i)
        : "r"(nb), "r"(thr), "r"(delta), "r"(A), "r"(B));
    const bool acc = acc_i != 0u;
    uint32_t nn;
    int dNb, dn;
    if constexpr (ALGO == ALGO_TAYLOR) {
        nn = nb + (uint32_t)delta;
        dNb = delta;
        dn = 1 - 2 * (int)(nb & 1u);
    } else {
        nn = nb ^ 1u;
        dNb = 1 - 2 * (int)nb;
        dn = dNb;
    }
    if (acc) {                                                  // :152-154
        bonds[off] = (uint8_t)nn;
        c.Nb += dNb;
        c.npar += dn;
    }
    c.head2 = head_new;
    c.closed = acc ? pc : closed;
}

// R = steps per round; P = R/2 producer lanes per chain (each draws one Philox block = two steps)
template <int G> struct LatCfg {
    static constexpr int R = (G >= 8) ? 16 : 2 * G;
    static constexpr int P = R / 2;
};

template <int ALGO, int G, bool HIST, bool LOWT>
__global__ void __launch_bounds__(32)
worm_lat_kernel(const PhiloxArgs a, const LatGeo g)
{
    constexpr int R = LatCfg<G>::R;
    constexpr int P = LatCfg<G>::P;
    constexpr int CH = 32 / G;
    extern __shared__ __align__(16) uint8_t smem8[];
    const int lane = threadIdx.x;
    const int sub = lane % G;            // lane within the chain's group
    const int q = lane / G;              // chain slot within the warp
    const int64_t cidx = (int64_t)blockIdx.x * CH + q;
    const bool live = cidx < a.n_chains;
    const bool writer = live && (sub == 0);
    const int64_t cl = live ? cidx : 0;  // dead groups shadow chain 0 in their own shared-memory slot, never write

    uint8_t *bonds = smem8;                                               // [CH][2N]
    uint4 *recA = reinterpret_cast<uint4 *>(smem8 + g.rec_off);           // [R][CH]
    uint4 *recB = recA + R * CH;                                          // [R][CH]
    uint32_t *hist = HIST ? reinterpret_cast<uint32_t *>(smem8 + g.hist_off) + q * g.N : nullptr;
    const int chain_off = q * g.nb2;

    const PhiloxChain prm = a.prm[cl];
    uint8_t *gb = a.bonds + cl * (int64_t)g.nb2;
    LatState c;
    c.head2 = 2 * a.head[cl];
    c.tail2 = 2 * a.tail[cl];
    c.closed = (c.head2 == c.tail2);
    const int64_t *ac0 = a.acc + cl * ACC_STRIDE;
    c.Z = (uint64_t)ac0[0]; c.SNb = (uint64_t)ac0[1]; c.SNb2 = (uint64_t)ac0[2];
    c.Sn = (uint64_t)ac0[3]; c.Sn2 = (uint64_t)ac0[4];
    // stage the chain's bonds (16-byte vectors, the group's lanes strided) and total them
    {
        int Nb = 0, npar = 0;
        const uint4 *src = reinterpret_cast<const uint4 *>(gb);
        uint4 *dst = reinterpret_cast<uint4 *>(bonds + chain_off);
        for (int i = sub; i < g.nb2 / 16; i += G) {
            const uint4 v = live ? src[i] : make_uint4(0, 0, 0, 0);
            dst[i] = v;
            Nb += (int)(__vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0));
            npar += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                    __popc(v.w & 0x01010101u);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            Nb += __shfl_xor_sync(0xffffffffu, Nb, o);
            npar += __shfl_xor_sync(0xffffffffu, npar, o);
        }
        c.Nb = Nb; c.npar = npar;
        if constexpr (HIST) {
            for (int i = sub; i < g.N; i += G) hist[i] = 0u;
        }
    }
    __syncwarp();
    const bool hist_writer = HIST && writer && a.hist != nullptr;
    int32_t nd = (a.n_dumps && live) ? a.n_dumps[cl] : 0;

    const uint64_t chain_id = a.chain_id0 + (uint64_t)cl;
    const uint32_t k0 = (uint32_t)prm.seed, k1 = (uint32_t)(prm.seed >> 32);
    const uint32_t c2 = (uint32_t)chain_id, c3 = (uint32_t)(chain_id >> 32);
    const int psub = sub < P ? sub : 0;                  // producer slot of this lane (lanes >= P idle along)

    // this lane's two records of the round whose base step is `pbase`
    uint4 na0, nb0, na1, nb1;
    uint64_t pbase = ~0ull;
    auto produce = [&](uint64_t base) {
        const uint64_t pair = (base >> 1) + (uint64_t)psub;
        const uint4 r = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), c2, c3, k0, k1);
        make_record<ALGO, LOWT>(g, prm, r.x, r.y, chain_off, na0, nb0);
        make_record<ALGO, LOWT>(g, prm, r.z, r.w, chain_off, na1, nb1);
    };
    auto publish = [&]() {
        if (sub < P) {
            recA[(2 * sub) * CH + q] = na0; recB[(2 * sub) * CH + q] = nb0;
            recA[(2 * sub + 1) * CH + q] = na1; recB[(2 * sub + 1) * CH + q] = nb1;
        }
    };

    uint64_t s = a.step_begin;
    while (s < a.step_end) {
        bool meas, dump_first;
        const uint64_t seg_end = next_segment(a, s, meas, dump_first);
        if (dump_first && live && c.closed) {                    // worm_ising_2d.cpp:113-124
            if (nd < a.max_dumps) {
                if (writer && a.events) {
                    int64_t *e = a.events + ((int64_t)cl * a.max_dumps + nd) * 3;
                    e[0] = (int64_t)s; e[1] = (int64_t)c.Z; e[2] = (int64_t)c.SNb;
                }
                if (a.dumps) {
                    uint4 *dst = reinterpret_cast<uint4 *>(a.dumps + ((int64_t)cl * a.max_dumps + nd) * g.nb2);
                    const uint4 *src = reinterpret_cast<const uint4 *>(bonds + chain_off);
                    for (int i = sub; i < g.nb2 / 16; i += G) dst[i] = src[i];
                }
            }
            ++nd;
        }
        while (s < seg_end) {
            const uint64_t s0 = s & ~1ull;                       // rounds start on a Philox pair
            if (pbase != s0) produce(s0);
            __syncwarp();
            publish();
            __syncwarp();
            if (s == s0 && seg_end - s0 >= (uint64_t)R) {
                // full round: records to registers, then R steps with no branch; the Philox block and
                // the records of the NEXT round are computed in the same basic block (overlap)
                uint4 RA[R], RB[R];
#pragma unroll
                for (int j = 0; j < R; ++j) { RA[j] = recA[j * CH + q]; RB[j] = recB[j * CH + q]; }
                pbase = s0 + (uint64_t)R;
                if (meas) {
                    produce(pbase);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if (j == 0) lat_step<ALGO, HIST, true, true, true>(g, bonds, hist, hist_writer, c, RA[j], RB[j], (int)RB[j + 1].x);
                        else if (j < R - 1) lat_step<ALGO, HIST, true, false, true>(g, bonds, hist, hist_writer, c, RA[j], RB[j], (int)RB[j + 1].x);
                        else lat_step<ALGO, HIST, true, false, false>(g, bonds, hist, hist_writer, c, RA[j], RB[j], 0);
                    }
                } else {
                    produce(pbase);
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        if (j == 0) lat_step<ALGO, HIST, false, true, true>(g, bonds, hist, hist_writer, c, RA[j], RB[j], (int)RB[j + 1].x);
                        else if (j < R - 1) lat_step<ALGO, HIST, false, false, true>(g, bonds, hist, hist_writer, c, RA[j], RB[j], (int)RB[j + 1].x);
                        else lat_step<ALGO, HIST, false, false, false>(g, bonds, hist, hist_writer, c, RA[j], RB[j], 0);
                    }
                }
                s += (uint64_t)R;
            } else {
                // partial round (segment edges): records straight from shared memory, no folding
                const int jb = (int)(s - s0);
                const uint64_t left = seg_end - s0;
                const int je = left < (uint64_t)R ? (int)left : R;
                for (int j = jb; j < je; ++j) {
                    const uint4 xa = recA[j * CH + q], xb = recB[j * CH + q];
                    if (meas) lat_step<ALGO, HIST, true, true, false>(g, bonds, hist, hist_writer, c, xa, xb, 0);
                    else lat_step<ALGO, HIST, false, true, false>(g, bonds, hist, hist_writer, c, xa, xb, 0);
                }
                s = s0 + (uint64_t)je;
                pbase = ~0ull;
            }
        }
    }

    // write back
    __syncwarp();
    if (live) {
        uint4 *dst = reinterpret_cast<uint4 *>(gb);
        const uint4 *src = reinterpret_cast<const uint4 *>(bonds + chain_off);
        for (int i = sub; i < g.nb2 / 16; i += G) dst[i] = src[i];
        if constexpr (HIST) {
            if (a.hist && a.group_index) {
                unsigned long long *row =
                    reinterpret_cast<unsigned long long *>(a.hist + (int64_t)a.group_index[cl] * g.N);
                for (int i = sub; i < g.N; i += G) {
                    const uint32_t v = hist[i];
                    if (v) atomicAdd(row + i, (unsigned long long)v);
                }
            }
        }
    }
    if (writer) {
        a.head[cl] = (c.head2 & g.M) >> 1;
        a.tail[cl] = (c.tail2 & g.M) >> 1;
        int64_t *ac = a.acc + cl * ACC_STRIDE;
        const int64_t iters = measured_iters(a);
        add_group_acc(a, cl, ac, c.Z, c.SNb, c.SNb2, c.Sn, c.Sn2, iters);
        ac[0] = (int64_t)c.Z; ac[1] = (int64_t)c.SNb; ac[2] = (int64_t)c.SNb2;
        ac[3] = (int64_t)c.Sn; ac[4] = (int64_t)c.Sn2;
        ac[5] += iters;
        if (a.n_dumps) a.n_dumps[cl] = nd;
    }
}

// ---------------------------------------------------------------------------------------------
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

template <int ALGO, int G, bool HIST, bool LOWT>
cudaError_t launch_lat(const PhiloxArgs &a, const LatGeo &g, size_t smem_bytes, cudaStream_t st)
{
    auto kern = worm_lat_kernel<ALGO, G, HIST, LOWT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    const int64_t blocks = (a.n_chains + g.ch - 1) / g.ch;
    kern<<<(unsigned)blocks, 32, smem_bytes, st>>>(a, g);
    return cudaGetLastError();
}

template <int ALGO, int G>
cudaError_t launch_lat_g(const PhiloxArgs &a, const LatGeo &g, bool hist, bool lowt, size_t smem_bytes, cudaStream_t st)
{
    if (hist) {
        if (lowt) return launch_lat<ALGO, G, true, true>(a, g, smem_bytes, st);
        return launch_lat<ALGO, G, true, false>(a, g, smem_bytes, st);
    }
    if (lowt) return launch_lat<ALGO, G, false, true>(a, g, smem_bytes, st);
    return launch_lat<ALGO, G, false, false>(a, g, smem_bytes, st);
}

template <int ALGO>
cudaError_t launch_lat_a(const PhiloxArgs &a, const LatGeo &g, int G, bool hist, bool lowt, size_t smem_bytes, cudaStream_t st)
{
    switch (G) {
    case 4: return launch_lat_g<ALGO, 4>(a, g, hist, lowt, smem_bytes, st);
    case 8: return launch_lat_g<ALGO, 8>(a, g, hist, lowt, smem_bytes, st);
    case 32: return launch_lat_g<ALGO, 32>(a, g, hist, lowt, smem_bytes, st);
    default: return cudaErrorInvalidValue;
    }
}

// shared memory of a latency-mode block with G lanes per chain (0 if L is not a power of two >= 4)
size_t lat_smem_bytes(int L, int G, bool hist, LatGeo *out)
{
    if (L < 4 || (L & (L - 1))) return 0;
    LatGeo g;
    g.L = L; g.N = L * L; g.nb2 = 2 * g.N; g.M = g.nb2 - 1; g.XM = 2 * L - 2;
    g.ch = 32 / G;
    const int R = (G >= 8) ? 16 : 2 * G;
    g.rec_off = g.ch * g.nb2;
    g.hist_off = g.rec_off + 2 * R * g.ch * 16;
    const size_t total = (size_t)g.hist_off + (hist ? (size_t)g.ch * g.N * 4 : 0);
    if (out) *out = g;
    return total;
}

}  // namespace

cudaError_t launch_philox(const PhiloxArgs &a, int lanes_per_chain, bool lowt, int smem_optin, int sm_count,
                          cudaStream_t st, const char **why)
{
    *why = "";
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    int G = lanes_per_chain;
    if (G == 0) {
        // Latency mode while the chains fit one warp per SM sub-partition (x4 for G = 4): among the
        // G whose shared memory fits, the largest one that still keeps one warp per sub-partition.
        G = 1;
        const int64_t slots = (int64_t)sm_count * 4;
        for (int cand : {4, 8, 32}) {
            const size_t need = lat_smem_bytes(a.L, cand, hist, nullptr);
            if (need == 0 || need > (size_t)smem_optin) continue;
            const int64_t warps = (a.n_chains + 32 / cand - 1) / (32 / cand);
            if (warps <= slots) G = cand;
            else if (G == 1 && cand == 4 && warps <= 4 * slots) G = 4;
        }
    }
    if (G == 4 || G == 8 || G == 32) {
        LatGeo g;
        const size_t need = lat_smem_bytes(a.L, G, hist, &g);
        if (need == 0) { *why = "latency mode (lanes_per_chain 4/8/32) needs a power-of-two L >= 4"; return cudaErrorInvalidValue; }
        if (need > (size_t)smem_optin) { *why = "lattice too large for latency mode with this lanes_per_chain"; return cudaErrorInvalidValue; }
        if (a.algo == ALGO_TAYLOR) return launch_lat_a<ALGO_TAYLOR>(a, g, G, hist, lowt, need, st);
        return launch_lat_a<ALGO_BINARY>(a, g, G, hist, lowt, need, st);
    }
    const bool force_global = (G == -1);
    if (G != 1 && G != -1) { *why = "lanes_per_chain must be 0, 1, 4, 8, 32 or -1"; return cudaErrorInvalidValue; }

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
