// worm_lat.cu -- production worm kernel, "latency mode" (sm_100a, hand-written CUDA + inline PTX).
//
// For sweeps with fewer Markov chains than the GPU has issue slots (BASELINE configs[1]: 4096
// chains on 148 x 4 SM sub-partitions).  A chain is a sequence of dependent steps, so
//     throughput = chains / (cycles per step),
// and the kernel is built around the per-step critical path
//     head --IADD,LOP3,IADD--> bond address --LDS--> occupation --ISETP--> accept --SEL--> head.
//
// Warp specialisation.  A block is 4 WALKER warps + 4 PRODUCER warps (warp w and warp 4+w form a
// pair that lands on the same SM sub-partition, so the hardware scheduler fills the walker's
// dependency stalls with the producer's independent instructions):
//   * a producer lane draws Philox blocks (two steps each) for the chains of its pair and turns
//     them, ahead of time, into ready-to-use step records: pre-decoded head increment and wrap
//     mask, absolute shared-memory address of the bond array (+ y bit), kick site, occupation
//     threshold.  Records go into a double-buffered shared-memory ring, one round (R steps per
//     chain) per buffer; pairs meet at a named barrier (bar.sync, 64 threads) once per round.
//   * a walker warp owns CH = 8, 4 or 1 chains, one per lane; its other lanes are predicated off
//     (SIMD lanes are free here, warp issue slots and shared-memory wavefronts are not: a 128-bit
//     load by 8 lanes is one wavefront, by 32 lanes four).  Records are fetched two steps ahead,
//     and the kick that follows a closing move is folded into the accept select, so the
//     head -> head dependency of consecutive steps is a single select after the accept test.
//     Walkers are the block's high warps: the sub-partition arbiter prefers the higher warp id.
// Power-of-two L only: the head is a doubled site index whose x and y bit fields wrap by masking.
//
// Algorithm and random stream: oracle/worm_oracle.c:worm_oracle_philox_run (bit-for-bit), i.e. the
// reference loop worm_ising_2d.cpp:111-157 on a Philox4x32-10 stream.
#include <type_traits>

#include "worm_run.cuh"

namespace worm {

namespace {

constexpr int NW = 4;                  // walker warps per block (= producer warps)

struct LatGeo {
    int L, N, nb2;       // nb2 = 2N bytes of bond storage per chain (uint8 per bond)
    int XM;              // 2L - 2: x field of a doubled site index
    int YM;              // (2N - 1) & ~(2L - 1): its y field
    int rec_off;         // byte offsets inside dynamic shared memory
    int hist_off;
};

template <int G> struct LatCfg {
    static constexpr int CH = 32 / G;                 // chains per walker warp
    static constexpr int BCH = NW * CH;               // chains per block
    static constexpr int U = (G == 4) ? 2 : 1;        // Philox blocks per producer lane per round
    static constexpr int R = 64 * U / CH;             // steps per chain per round
    static constexpr int PR = R / 2;                  // Philox blocks (pairs of steps) per chain per round
    static constexpr int BUF = R * CH * 16;           // bytes of one buffer of A (or B) records
    static constexpr int REC_B = 2 * BUF;             // byte offset of the B records (two buffers of A first)
    static constexpr int RECW = 4 * BUF;              // record bytes per walker warp
};

// ---- shared memory by absolute 32-bit shared-window address.  All hot-loop accesses are volatile
// asm with a memory clobber: program order is the order written here.
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void pair_barrier(int id)
{
    asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory");
}

struct LatState {
    int head2, tail2;    // doubled site indices in [0, 2N)
    uint32_t closed;     // 0 / 1.  When 1, head2 already holds the (folded) kick of the coming step.
    int Nb, npar;
    uint32_t z32, snb32, sn32;           // partial sums, flushed once per round
    uint64_t Z, SNb, SNb2, Sn, Sn2;
};

// A step record: everything about step s that does not depend on the chain's state.
//   ra = {add, keep, addneg, base}:  nh2  = (head2 & keep) | ((head2 + add) & ~keep)      new head
//                                    own2 = (head2 & keep) | ((head2 + addneg) & ~keep)   site owning the bond
//                                    bond address = own2 + base      (base = chain's bonds + y bit)
//   rb = {kick2, thr, delta, -}:     accept iff (nb < thr) != (delta < 0);  Taylor: nb += delta
template <int ALGO, bool LOWT>
__device__ __forceinline__ void make_record(const LatGeo &g, const PhiloxChain &prm, uint32_t a, uint32_t b,
                                            uint32_t bonds_abs, uint4 &ra, uint4 &rb)
{
    const uint32_t d = b >> 30;
    const bool ym = d & 1u, neg = d & 2u;
    const int stepsz = ym ? 2 * g.L : 2;
    const int add = neg ? -stepsz : stepsz;
    ra.x = (uint32_t)add;
    ra.y = ym ? ~(uint32_t)g.YM : ~(uint32_t)g.XM;
    ra.z = neg ? (uint32_t)add : 0u;
    ra.w = bonds_abs + (uint32_t)ym;
    rb.x = 2u * __umulhi(b << 3, (uint32_t)g.N);                 // kick site :128
    if constexpr (ALGO == ALGO_TAYLOR) {
        const bool dec = (b >> 29) & 1u;                         // :141
        uint32_t qi, qd;
        if constexpr (!LOWT) {
            // T >= 1: kfix - 1 < 2^32 and tfix >= 2^32 > a.
            // (nb+1)*a < kfix  <=>  nb < floor((kfix-1)/a); a == 0 accepts up to the storage cap.
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

// One worm step on a pre-decoded record.  On entry head2 already holds this step's kick when the
// worm is closed; on exit it holds the NEXT step's kick (kick2_next) when the worm is closed then.
template <int ALGO, bool HIST, bool MEAS>
__device__ __forceinline__ void lat_step(const LatGeo &g, uint32_t hist_abs, uint32_t hist_writer,
                                         LatState &c, const uint4 ra, const uint4 rb, const int kick2_next)
{
    const int add = (int)ra.x, keep = (int)ra.y, addneg = (int)ra.z;
    const uint32_t base = ra.w;
    const int kick2 = (int)rb.x, delta = (int)rb.z;
    const uint32_t thr = rb.y;
    const bool closed = c.closed != 0u;
    if constexpr (MEAS) {
        if constexpr (HIST) {
            // G(r) bin of (head - tail) before the kick; a closed worm is bin 0
            const int t1 = c.head2 - c.tail2;
            const int t2 = (c.head2 | g.XM | 1) - (c.tail2 & g.YM);            // y fields, no borrow from x
            const int bin4 = closed ? 0 : (((t1 & g.XM) | (t2 & g.YM)) << 1);  // byte offset of the u32 bin
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}"
                         :: "r"(hist_abs + (uint32_t)bin4), "r"(hist_writer) : "memory");
        }
        if (closed) {                                           // worm_ising_2d.cpp:130-132
            c.z32 += 1u;
            c.snb32 += (uint32_t)c.Nb;
            c.sn32 += (uint32_t)c.npar;
            c.SNb2 += (uint64_t)(uint32_t)c.Nb * (uint32_t)c.Nb;
            c.Sn2 += (uint64_t)(uint32_t)c.npar * (uint32_t)c.npar;
        }
    }
    const int tail2 = closed ? kick2 : c.tail2;                 // kick :128-129 (head2 has it already)
    c.tail2 = tail2;
    const int head2 = c.head2;
    const int t = head2 + add;
    const int nh2 = (head2 & keep) | (t & ~keep);               // neighbour table :88-95
    const int u = head2 + addneg;
    const int own = (head2 & keep) | (u & ~keep);               // bond_number :214-244
    const uint32_t addr = (uint32_t)own + base;
    const uint32_t nb = lds_u8(addr);                           // :139
    const uint32_t pc = (nh2 == tail2) ? 1u : 0u;               // closed after an accepted move?
    const int A = pc ? kick2_next : nh2;
    const int B = closed ? kick2_next : head2;
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
    // accept test (:151), head select and bond store (:152-154) in one opaque block: the compiler
    // must not re-associate "closed ? kick : ..." across the select (that puts two more operations
    // on the head -> head critical path)
    int head_new;
    uint32_t acc_i;
    asm volatile("{\n\t.reg .pred pa, pd;\n\t"
                 "setp.lt.u32 pa, %2, %3;\n\t"
                 "setp.lt.s32 pd, %4, 0;\n\t"
                 "xor.pred pa, pa, pd;\n\t"
                 "selp.b32 %0, %5, %6, pa;\n\t"
                 "selp.u32 %1, 1, 0, pa;\n\t"
                 "@pa st.shared.u8 [%7], %8;\n\t}"
                 : "=r"(head_new), "=r"(acc_i)
                 : "r"(nb), "r"(thr), "r"(delta), "r"(A), "r"(B), "r"(addr), "r"(nn)
                 : "memory");
    const bool acc = acc_i != 0u;
    if (acc) {
        c.Nb += dNb;
        c.npar += dn;
    }
    c.head2 = head_new;
    c.closed = acc ? pc : c.closed;
}

__device__ __forceinline__ void flush_partials(LatState &c)
{
    c.Z += c.z32; c.SNb += c.snb32; c.Sn += c.sn32;
    c.z32 = c.snb32 = c.sn32 = 0u;
}

// The round schedule both roles follow: rounds start on a Philox pair, never straddle a segment
// (MEAS change / dump multiple), and hold at most R steps.
struct Round {
    uint64_t s0;     // base step of the record buffer (even)
    int jb, je;      // steps [s0 + jb, s0 + je) are walked in this round
};
template <int R>
__device__ __forceinline__ Round make_round(uint64_t s, uint64_t seg_end)
{
    Round r;
    r.s0 = s & ~1ull;
    r.jb = (int)(s - r.s0);
    const uint64_t left = seg_end - r.s0;
    r.je = left < (uint64_t)R ? (int)left : R;
    return r;
}

template <int ALGO, int G, bool HIST, bool LOWT>
__global__ void __launch_bounds__(2 * NW * 32)
worm_lat_kernel(const PhiloxArgs a, const LatGeo g)
{
    using Cfg = LatCfg<G>;
    constexpr int R = Cfg::R, CH = Cfg::CH, BCH = Cfg::BCH, PR = Cfg::PR, U = Cfg::U;
    extern __shared__ __align__(16) uint8_t smem8[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int w = warp & (NW - 1);                       // pair index: warps w and NW + w share a sub-partition
    const bool is_producer = warp < NW;                  // walkers are the high warps (arbiter priority)
    const int64_t chain0 = (int64_t)blockIdx.x * BCH;    // first chain of the block
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem8);
    const uint32_t rec_w = sbase + (uint32_t)g.rec_off + (uint32_t)(w * Cfg::RECW);
    const int bar_id = 1 + w;

    // ---- stage the block's bonds (all 8 warps, 16-byte vectors)
    {
        const int vec_per_chain = g.nb2 / 16;
        const int total = BCH * vec_per_chain;
        uint4 *dst = reinterpret_cast<uint4 *>(smem8);
        for (int i = tid; i < total; i += 2 * NW * 32) {
            const int cslot = i / vec_per_chain;
            const int64_t cg = chain0 + cslot;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cg < a.n_chains)
                v = reinterpret_cast<const uint4 *>(a.bonds + cg * (int64_t)g.nb2)[i - cslot * vec_per_chain];
            dst[i] = v;
        }
        if constexpr (HIST) {
            uint32_t *h = reinterpret_cast<uint32_t *>(smem8 + g.hist_off);
            for (int i = tid; i < BCH * g.N; i += 2 * NW * 32) h[i] = 0u;
        }
    }
    __syncthreads();

    if (is_producer) {
        // ================================ producer warp ================================
        // Philox block b = u*32 + lane of a round belongs to chain slot q = b % CH, pair p = b / CH:
        // consecutive lanes store consecutive 16-byte records (conflict-free 128-bit stores).
        PhiloxChain prm[U];
        uint32_t c2[U], c3[U], bonds_abs[U], rec_q[U];
        int p_of[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = u * 32 + lane;
            const int q = b % CH;
            p_of[u] = b / CH;
            const int64_t cg = chain0 + w * CH + q;
            const int64_t cl = cg < a.n_chains ? cg : 0;
            prm[u] = a.prm[cl];
            const uint64_t chain_id = a.chain_id0 + (uint64_t)cl;
            c2[u] = (uint32_t)chain_id; c3[u] = (uint32_t)(chain_id >> 32);
            bonds_abs[u] = sbase + (uint32_t)((w * CH + q) * g.nb2);
            rec_q[u] = rec_w + (uint32_t)q * 16u + (uint32_t)(2 * p_of[u]) * (CH * 16);
        }
        uint32_t buf = 0u;
        uint64_t s = a.step_begin;
        while (s < a.step_end) {
            bool meas, dump_first;
            const uint64_t seg_end = next_segment(a, s, meas, dump_first);
            while (s < seg_end) {
                const Round rd = make_round<R>(s, seg_end);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t pair = (rd.s0 >> 1) + (uint64_t)p_of[u];
                    const uint4 r = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), c2[u], c3[u],
                                                  (uint32_t)prm[u].seed, (uint32_t)(prm[u].seed >> 32));
                    uint4 ra0, rb0, ra1, rb1;
                    make_record<ALGO, LOWT>(g, prm[u], r.x, r.y, bonds_abs[u], ra0, rb0);
                    make_record<ALGO, LOWT>(g, prm[u], r.z, r.w, bonds_abs[u], ra1, rb1);
                    const uint32_t p0 = rec_q[u] + buf;
                    sts_v4(p0, ra0); sts_v4(p0 + Cfg::REC_B, rb0);
                    sts_v4(p0 + CH * 16, ra1); sts_v4(p0 + CH * 16 + Cfg::REC_B, rb1);
                }
                pair_barrier(bar_id);
                buf ^= (uint32_t)Cfg::BUF;
                s = rd.s0 + (uint64_t)rd.je;
            }
        }
        pair_barrier(bar_id);                            // matches the walker's barrier before its last step
    } else {
        // ================================= walker warp =================================
        // One lane per chain; the other lanes of the warp retire here (barriers count warps, and a
        // shared-memory access by CH lanes costs CH/8 wavefronts instead of 4).
        if (lane >= CH) return;
        constexpr unsigned wmask = (CH == 32) ? 0xffffffffu : ((1u << CH) - 1u);
        const int q = lane;                  // chain slot within the warp
        const int cslot = w * CH + q;
        const int64_t cidx = chain0 + cslot;
        const bool live = cidx < a.n_chains;
        const int64_t cl = live ? cidx : 0;  // dead slots shadow chain 0 in their own shared-memory slot, never write
        const uint8_t *bonds = smem8 + cslot * g.nb2;
        uint32_t rec_q = rec_w + (uint32_t)q * 16u;
        uint32_t hist_abs = sbase + (uint32_t)g.hist_off + (uint32_t)(cslot * g.N) * 4u;
        const uint32_t hist_writer = (HIST && live && a.hist != nullptr) ? 1u : 0u;
        asm volatile("" : "+r"(rec_q), "+r"(hist_abs));                  // keep them in registers

        LatState c;
        c.head2 = 2 * a.head[cl];
        c.tail2 = 2 * a.tail[cl];
        c.closed = (c.head2 == c.tail2) ? 1u : 0u;
        c.z32 = c.snb32 = c.sn32 = 0u;
        const int64_t *ac0 = a.acc + cl * ACC_STRIDE;
        c.Z = (uint64_t)ac0[0]; c.SNb = (uint64_t)ac0[1]; c.SNb2 = (uint64_t)ac0[2];
        c.Sn = (uint64_t)ac0[3]; c.Sn2 = (uint64_t)ac0[4];
        {
            int Nb = 0, npar = 0;
            const uint4 *src = reinterpret_cast<const uint4 *>(bonds);
            for (int i = 0; i < g.nb2 / 16; ++i) {
                const uint4 v = src[i];
                Nb += (int)(__vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0));
                npar += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                        __popc(v.w & 0x01010101u);
            }
            c.Nb = Nb; c.npar = npar;
        }
        int32_t nd = (a.n_dumps && live) ? a.n_dumps[cl] : 0;

        auto rec_addr = [&](uint32_t buf, int j) { return rec_q + buf + (uint32_t)(j * CH * 16); };

        // A full round (steps 0..R-1 of the buffer at `cur`), software-pipelined two steps deep: on
        // entry (p0a,p0b) / (p1a,p1b) hold the records of steps 0 / 1; step j fetches the record of
        // step j+2, which for the last two steps lives in the other buffer (complete once the pair
        // barrier has been passed).  On exit they hold the records of steps 0 / 1 of the next round.
        uint4 p0a, p0b, p1a, p1b;
        auto full_round = [&](auto meas_tag, uint32_t cur, uint32_t nxt) {
            constexpr bool MEAS = decltype(meas_tag)::value;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (j == R - 2) pair_barrier(bar_id);
                const uint32_t nrec = (j < R - 2) ? rec_addr(cur, j + 2) : rec_addr(nxt, j + 2 - R);
                const uint4 fa = lds_v4(nrec), fb = lds_v4(nrec + Cfg::REC_B);
                lat_step<ALGO, HIST, MEAS>(g, hist_abs, hist_writer, c, p0a, p0b, (int)p1b.x);
                p0a = p1a; p0b = p1b; p1a = fa; p1b = fb;
            }
            if constexpr (MEAS) flush_partials(c);
        };
        // Any other round (segment edges): steps [jb, je), records straight from shared memory.  The
        // kick folded into the last step is the first record of the next round (slot jbn of `nxt`).
        auto edge_round = [&](auto meas_tag, uint32_t cur, uint32_t nxt, int jb, int je, int jbn) {
            constexpr bool MEAS = decltype(meas_tag)::value;
            for (int j = jb; j < je; ++j) {
                const uint32_t rec = rec_addr(cur, j);
                const uint4 xa = lds_v4(rec), xb = lds_v4(rec + Cfg::REC_B);
                uint32_t nrec = rec_addr(cur, j + 1);
                if (j == je - 1) { pair_barrier(bar_id); nrec = rec_addr(nxt, jbn); }
                const uint4 nb_ = lds_v4(nrec + Cfg::REC_B);
                lat_step<ALGO, HIST, MEAS>(g, hist_abs, hist_writer, c, xa, xb, (int)nb_.x);
            }
            flush_partials(c);
        };

        uint32_t cur = 0u;
        uint64_t s = a.step_begin;
        bool first = true;
        bool piped = false;                  // (p0, p1) hold the first two records of the coming round
        while (s < a.step_end) {
            bool meas, dump_first;
            const uint64_t seg_end = next_segment(a, s, meas, dump_first);
            if (dump_first) {                                        // worm_ising_2d.cpp:113-124
                // the whole warp copies the bonds of every chain that dumps now
                const bool want = live && c.closed;
                const bool store = want && nd < a.max_dumps;
                if (store && a.events) {
                    int64_t *e = a.events + ((int64_t)cl * a.max_dumps + nd) * 3;
                    e[0] = (int64_t)s; e[1] = (int64_t)c.Z; e[2] = (int64_t)c.SNb;
                }
                unsigned dm = __ballot_sync(wmask, store && a.dumps != nullptr);
                while (dm) {
                    const int src_lane = __ffs(dm) - 1;
                    dm &= dm - 1;
                    const long long chain = __shfl_sync(wmask, (long long)cl, src_lane);
                    const int slot = __shfl_sync(wmask, nd, src_lane);
                    uint4 *dst = reinterpret_cast<uint4 *>(a.dumps + (chain * a.max_dumps + slot) * g.nb2);
                    const uint4 *src = reinterpret_cast<const uint4 *>(smem8 + (w * CH + src_lane) * g.nb2);
                    for (int i = lane; i < g.nb2 / 16; i += CH) dst[i] = src[i];
                }
                if (want) ++nd;
            }
            while (s < seg_end) {
                const Round rd = make_round<R>(s, seg_end);
                const uint64_t e = rd.s0 + (uint64_t)rd.je;          // first step of the next round
                const uint32_t nxt = cur ^ (uint32_t)Cfg::BUF;
                const bool full = (rd.jb == 0 && rd.je == R);
                if (first) {
                    pair_barrier(bar_id);                            // the first round's records are complete
                    const uint4 kb = lds_v4(rec_addr(cur, rd.jb) + Cfg::REC_B);
                    c.head2 = c.closed ? (int)kb.x : c.head2;        // kick of the first step
                    first = false;
                }
                if (full) {
                    if (!piped) {
                        p0a = lds_v4(rec_addr(cur, 0)); p0b = lds_v4(rec_addr(cur, 0) + Cfg::REC_B);
                        p1a = lds_v4(rec_addr(cur, 1)); p1b = lds_v4(rec_addr(cur, 1) + Cfg::REC_B);
                    }
                    if (meas) full_round(std::true_type{}, cur, nxt);
                    else full_round(std::false_type{}, cur, nxt);
                    piped = true;
                } else {
                    if (meas) edge_round(std::true_type{}, cur, nxt, rd.jb, rd.je, (int)(e & 1ull));
                    else edge_round(std::false_type{}, cur, nxt, rd.jb, rd.je, (int)(e & 1ull));
                    piped = false;
                }
                cur = nxt;
                s = e;
            }
        }

        if (first) pair_barrier(bar_id);                             // empty step range: meet the producer once
        if (live) {
            // a closed worm sits on its tail (head2 holds the folded kick of the step after the last)
            a.head[cl] = (c.closed ? c.tail2 : c.head2) >> 1;
            a.tail[cl] = c.tail2 >> 1;
            int64_t *ac = a.acc + cl * ACC_STRIDE;
            const int64_t iters = measured_iters(a);
            add_group_acc(a, cl, ac, c.Z, c.SNb, c.SNb2, c.Sn, c.Sn2, iters);
            ac[0] = (int64_t)c.Z; ac[1] = (int64_t)c.SNb; ac[2] = (int64_t)c.SNb2;
            ac[3] = (int64_t)c.Sn; ac[4] = (int64_t)c.Sn2;
            ac[5] += iters;
            if (a.n_dumps) a.n_dumps[cl] = nd;
        }
    }

    // ---- write back (the producer warps: the walkers' spare lanes have retired)
    __syncthreads();
    if (is_producer) {
        const int vec_per_chain = g.nb2 / 16;
        const int total = BCH * vec_per_chain;
        const uint4 *src = reinterpret_cast<const uint4 *>(smem8);
        for (int i = tid; i < total; i += NW * 32) {
            const int cslot = i / vec_per_chain;
            const int64_t cg = chain0 + cslot;
            if (cg < a.n_chains)
                reinterpret_cast<uint4 *>(a.bonds + cg * (int64_t)g.nb2)[i - cslot * vec_per_chain] = src[i];
        }
        if constexpr (HIST) {
            if (a.hist && a.group_index) {
                const uint32_t *h = reinterpret_cast<const uint32_t *>(smem8 + g.hist_off);
                for (int i = tid; i < BCH * g.N; i += NW * 32) {
                    const int cslot = i / g.N;
                    const int64_t cg = chain0 + cslot;
                    const uint32_t v = h[i];
                    if (cg < a.n_chains && v) {
                        unsigned long long *row =
                            reinterpret_cast<unsigned long long *>(a.hist + (int64_t)a.group_index[cg] * g.N);
                        atomicAdd(row + (i - cslot * g.N), (unsigned long long)v);
                    }
                }
            }
        }
    }
}

// shared memory layout of a latency-mode block with G lanes per chain (false if L is not a power of two >= 4)
bool lat_geometry(int L, int G, bool hist, LatGeo &g, size_t &total)
{
    if (L < 4 || (L & (L - 1))) return false;
    g.L = L; g.N = L * L; g.nb2 = 2 * g.N;
    g.XM = 2 * L - 2;
    g.YM = (g.nb2 - 1) & ~(2 * L - 1);
    const int ch = 32 / G;
    const int bch = NW * ch;
    const int U = (G == 4) ? 2 : 1;
    const int R = 64 * U / ch;
    g.rec_off = bch * g.nb2;
    g.hist_off = g.rec_off + NW * 4 * R * ch * 16;       // per pair: two buffers x (A, B) records
    total = (size_t)g.hist_off + (hist ? (size_t)bch * g.N * 4 : 0);
    return true;
}

template <int ALGO, int G, bool HIST, bool LOWT>
cudaError_t launch_lat(const PhiloxArgs &a, const LatGeo &g, size_t smem_bytes, cudaStream_t st)
{
    auto kern = worm_lat_kernel<ALGO, G, HIST, LOWT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    constexpr int BCH = LatCfg<G>::BCH;
    const int64_t blocks = (a.n_chains + BCH - 1) / BCH;
    kern<<<(unsigned)blocks, 2 * NW * 32, smem_bytes, st>>>(a, g);
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

}  // namespace

int lat_chains_per_block(int G) { return NW * (32 / G); }

size_t lat_smem_bytes(int L, int G, bool hist)
{
    LatGeo g;
    size_t total = 0;
    if (G != 4 && G != 8 && G != 32) return 0;
    return lat_geometry(L, G, hist, g, total) ? total : 0;
}

cudaError_t launch_philox_lat(const PhiloxArgs &a, int G, bool lowt, cudaStream_t st)
{
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    LatGeo g;
    size_t total = 0;
    if (!lat_geometry(a.L, G, hist, g, total)) return cudaErrorInvalidValue;
    if (a.algo == ALGO_TAYLOR) return launch_lat_a<ALGO_TAYLOR>(a, g, G, hist, lowt, total, st);
    return launch_lat_a<ALGO_BINARY>(a, g, G, hist, lowt, total, st);
}

}  // namespace worm
