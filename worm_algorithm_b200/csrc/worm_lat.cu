// worm_lat.cu -- production worm kernel, "latency mode" (sm_100a, hand-written CUDA + inline PTX).
//
// A Markov chain is a sequence of dependent steps, so
//     throughput = (chains in flight) / (cycles per step),
// and BASELINE configs[1] has only 4096 chains for 148 x 4 SM sub-partitions.  The kernel is built
// around the cycles one step costs the warp that walks the chains: the dependency chain
//     head --IADD--> bond address --LDS--> occupation --ISETP--> accept --SEL--> head
// and the number of instructions that warp has to issue (a lone warp issues in order).
//
// Warp specialisation.  The warps of a block form groups of three roles:
//   * PRODUCER warps draw Philox blocks (two steps each, key schedule in registers) for the chains
//     of their group and turn them, ahead of time, into ready-to-use step records: pre-decoded head
//     increment and wrap mask, absolute shared-memory addresses of the two copies of the bond, kick
//     site, occupation threshold and accept polarity.  Records go into a double-buffered
//     shared-memory ring, one round (R steps per chain) per buffer.
//   * the WALKER warp owns CH chains, one per lane (lanes beyond CH retire at once).  Records are
//     fetched two steps ahead; the kick that follows a closing move is folded into the accept
//     select; everything independent of the loaded occupation is forced into the load's shadow.
//     It keeps no observables: it logs one byte per step (accepted / decrement / old parity / direction, and whether
//     the worm was closed at the start of the step), four steps to a word.
//   * ACCOUNTANT warps replay those logs two rounds later and accumulate Z, sum Nb, sum Nb^2,
//     sum n, sum n^2 with the exact Nb / n of every closed step; they also write the dump events
//     and, where the walker-side G(r) bins do not fit shared memory, the histogram: the accept log carries the move's
//     direction, so the displacement head - tail is replayed like Nb and runs of equal bins are added with red.global
//     (in solo mode at L <= 32 the walker itself counts in 16-bit shared-memory bins and sweeps them into HBM).
//   The group meets at one named barrier per round (16 steps per chain; 64 with one chain per walker).
//   The walker is a LONE in-order warp: everything outside its unrolled round is cold code for it (~45 cycles for every
//   instruction line it has to fetch, ~30 for every taken branch).  Hence one loop per MEAS value with the test at the
//   bottom, and for the dump rule (worm_ising_2d.cpp:113-125, the step range is cut at every multiple of write_steps)
//   tight steady-state loops over whole dump intervals in all three roles: pause, the interval's short round (first,
//   built by the producer lanes it needs, walked by a pipelined rolled loop), its full rounds.
//   CH = 32 ("solo"): one group per block -- the walker has an SM sub-partition to itself, 8 producers
//   and 4 accountants share the other three.  CH = 8 / 4 / 1: up to four
//   groups per block, the warps of a group on the same sub-partition (walker = highest warp id,
//   which the sub-partition arbiter prefers).
// Bond storage ("dual" layout): one 32-bit word per SITE holding the occupations of its four
// bonds (+x, +y, -x, -y), words interleaved across the chains of the walker.  A bond is stored at
// both of its sites, so the load address is head + constant (one IADD, no neighbour arithmetic in
// front of the load), a chain only ever touches its own banks (no conflicts between the chains of
// a warp), and an accepted move stores twice off the critical path.  Lattices whose dual layout
// does not fit use the compact one-byte-per-bond layout (one chain per walker) or, for the Taylor chain, the
// nibble layout: ONE BYTE PER SITE (+x occupation in the low nibble, +y in the high one; occupations 0..15, an
// occupation that reaches 16 is reported in WORM_ACC_STATUS like in worm_pack.cu), the bytes of a chain in 4-byte
// columns, one column (one bank) per chain of the walker (nib_fmt below).  A quarter of the dual layout's bytes: 32 chains per SM at L = 64 instead of 12,
// 12 at L = 128 instead of 4, 3 at L = 256 instead of 1 -- chains in flight are what bounds a large lattice.
// Power-of-two L only: the head is a scaled site index whose x and y bit fields wrap by masking.
//
// Algorithm and random stream: oracle/worm_oracle.c:worm_oracle_philox_run (bit-for-bit), i.e. the
// reference loop worm_ising_2d.cpp:111-157 on a Philox4x32-10 stream.
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include "worm_run.cuh"

namespace worm {

namespace {

enum { LAY_DUAL = 0, LAY_BYTE = 1, LAY_NIB = 2 };

struct LatGeo {
    int L, N, nb2;       // nb2 = 2N bonds per chain
    int nw;              // groups (walkers) per block
    int SC;              // bytes per unit of the scaled site index (dual: 4 * CH, compact: 2)
    int XM, YM;          // x / y bit fields of a scaled site index
    int XP, XN;          // what a +x / -x move adds to a scaled site index (nibble layout: +x also fills the gap of the x field)
    int region;          // bytes of bond storage per walker warp
    int rec_off;         // byte offset of the record rings inside dynamic shared memory
    int log_off;         // byte offset of the step logs
    int hist_off;        // byte offset of the 16-bit G(r) bins (solo mode with histograms)
    int htab_off;        // ... and of the four displacement-update entries {add, keep, gap, -}, one per direction
    int HXF, HYF;        // x / y bit fields of a displacement in bin-address format (solo mode with histograms)
    int bad_off;         // nibble layout: byte offset of the per-chain "input does not fit 4 bits" flags
    int lowt;            // some chain has T < 1 (general record builder)
    int zero;            // 0 (a value the compiler cannot see through)
};

// SH: per-chain G(r) bins in shared memory (solo mode with histograms), counted by the walker.
template <int CH, bool SH = false> struct LatCfg {
    static constexpr int NP = (CH == 32) ? 8 : (CH == 8) ? 2 : 1;   // producer warps per walker
    static constexpr int RPL = 2;                     // step records a producer lane builds per round (one Philox block holds two steps)
    static constexpr int ACH = (CH > 8) ? 8 : CH;     // chains per accountant warp
    static constexpr int NA = CH / ACH;               // accountant warps per walker
    static constexpr int GW = NP + NA + 1;            // warps of a group: producers, accountants, walker
    static constexpr int R = 32 * NP * RPL / CH;      // steps per chain per round
    static constexpr int BUF = R * CH * 16;           // bytes of one buffer of A (or B) records
    static constexpr int REC_B = 2 * BUF;             // byte offset of the B records (two buffers of A first)
    static constexpr int RECW = 4 * BUF;              // record bytes per walker warp
    static constexpr int LSTR = R + 4;                // pitch of a chain's log row (5 words: 32 lanes hit 32 banks)
    static constexpr int LOGH = LSTR * CH;            // bytes of one round's log (one byte per step and chain)
    static constexpr int LOGB = LOGH;
    static constexpr int LOGW = 3 * LOGB;             // three rounds in flight per walker warp
    static constexpr int PARTS = 32 / ACH;            // accountant lanes per chain
    static constexpr int SP = R / PARTS;              // steps per accountant lane per round
    static constexpr int FLUSH = 2048;                // SH: rounds between the walker's flushes of the 16-bit bins (FLUSH * R < 65536)
};

// ---- shared memory by absolute 32-bit shared-window address.  All hot-loop accesses are volatile
// asm with a memory clobber: program order is the order written here.
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
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
// one named barrier per walker group: its producers, its accountant and the walker (GW warps)
template <int GW>
__device__ __forceinline__ void group_barrier(int id)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(32 * GW) : "memory");
}

template <int... Js, class F>
__device__ __forceinline__ void unroll_steps(std::integer_sequence<int, Js...>, F &&f)
{
    (f(std::integral_constant<int, Js>{}), ...);
}

struct LatState {
    int head, tail;      // scaled site indices (site * SC)
    uint32_t closed;     // 0 / 1.  When 1, head already holds the (folded) kick of the coming step.
    uint32_t dh;         // solo mode with histograms: head - tail in bin-address format (0 iff closed)
    uint32_t aword;          // full rounds: log bytes of up to four steps, stored as one word
    uint32_t addr, nb;   // dual layout: shared address of the coming step's bond and its occupation, loaded
                         // at the end of the previous step (right behind that step's bond stores)
};

// Observables are NOT accumulated by the walker (every instruction of its dependent chain costs
// ~3 cycles of the step): it logs one byte per accepted step and one per closed step, and the
// producer warp of the pair replays the log two rounds later (account_round below).
//   accept log byte: 0x80 | (decrement ? 0x40 : 0) | (old occupation & 1, in bit 0 -- bit 4 when the occupation is
//                    the high nibble of a site byte);  0 = rejected
//                    bit 1: the worm was closed at the start of the step (set whether or not the move is accepted)
struct Account {
    int Nb, npar;                        // state at the start of the next round to be replayed (all lanes of a chain)
    uint64_t Z, SNb, SNb2, Sn, Sn2;      // this lane's share of the sums
    int32_t nd;                          // dump events seen (mirrors the walker's counter)
    int dx, dy;                          // G(r): head - tail (mod L) at the start of the next round to be replayed
    unsigned long long *hrow;            // G(r): the chain's row of the int64 histogram (null: do not count)
};

// A step record: everything about step s that does not depend on the chain's state.
//   ra = {add, keep, p, q}:  new head  nh = (head & keep) | ((head + add) & ~keep)
//        dual:    p = address of slot d of site 0, q = address of slot d^2 of site 0:
//                 the bond is read at head + p and written at head + p and nh + q
//        compact: p = addneg, q = address of bond byte (y bit) of site 0: the bond is owned by
//                 own = (head & keep) | ((head + addneg) & ~keep) and lives at own + q
//   rb = {kick, thr, delta, logbits}:  accept iff (nb < thr) xor (delta < 0);  Taylor: nb += delta (+1 / -1)
//   rc = {addh, keeph, gaph, -} (solo mode with histograms; a function of the direction only, so it is not a record
//        word but one of four table entries the walker looks up by the direction bits of rb.w -- a third record
//        word halved the round to fit beside the 64 KiB of bins): the same move applied to the displacement
//        head - tail kept in bin-address format,  dh' = (dh & keeph) | (((dh | gaph) + addh) & ~keeph).
//        Bin s = dy * L + dx of a chain lives at (s >> 1) * 128 + (s & 1) * 2 of the chain's bank, so
//        the x field is bit 1 plus bits 7.. : a +x move fills the gap (bits 2..6) with ones so that the
//        carry out of bit 1 reaches bit 7; a -x move borrows through the gap's zeros.
// numf = float(kfix - 1) * (1 + 2^-18), rounded up (see the threshold computation below)
// Nibble layout: a chain's site bytes are kept in 4-byte columns, chain slot q of a walker in column q, so that lane q
// only ever touches bank q (bytes interleaved chain by chain cost the solo walker 22 % of its cycles in bank
// conflicts).  Site s of slot q lives at region + nib_fmt(s) + 4 q with nib_fmt(s) = (s >> 2) * 4 CH + (s & 3): the
// scaled site index the walker carries.  Its x field is bits 0-1 plus the bits above the column index; a +x move adds
// 1 + (the gap filled with ones), so that the carry out of bit 1 crosses the gap, a -x move borrows through the gap's zeros,
// and the masked merge nh = (head & keep) | ((head + add) & ~keep) drops whatever the gap holds afterwards.
template <int CH>
__device__ __forceinline__ int nib_fmt(int s) { return (s >> 2) * (4 * CH) + (s & 3); }
template <int CH>
__device__ __forceinline__ int nib_site(int h) { return (h / (4 * CH)) * 4 + (h & 3); }

// the displacement update of direction d (bits 31-30 of the step's second random word)
__device__ __forceinline__ uint4 hist_entry(const LatGeo &g, uint32_t d)
{
    const bool ym = d & 1u, neg = d & 2u;
    const int hstep = ym ? g.L * 64 : 2;
    return make_uint4((uint32_t)(neg ? -hstep : hstep), ym ? ~(uint32_t)g.HYF : ~(uint32_t)g.HXF, (!ym && !neg) ? 0x7cu : 0u, 0u);
}

template <int ALGO, int LAY, int CH>
__device__ __forceinline__ void make_record(const LatGeo &g, const PhiloxChain &prm, const float numf, uint32_t a,
                                            uint32_t b, uint32_t chain_abs, uint4 &ra, uint4 &rb)
{
    const uint32_t d = b >> 30;                                  // :137
    const bool ym = d & 1u, neg = d & 2u;
    const int ystep = g.L * g.SC;
    const int add = ym ? (neg ? -ystep : ystep) : (neg ? g.XN : g.XP);
    ra.x = (uint32_t)add;
    ra.y = ym ? ~(uint32_t)g.YM : ~(uint32_t)g.XM;
    constexpr bool DUAL = (LAY == LAY_DUAL);
    const uint32_t sh4 = (LAY == LAY_NIB && ym) ? 4u : 0u;      // nibble layout: the +y occupation is the high nibble
    if constexpr (DUAL) {
        ra.z = chain_abs + d;
        ra.w = chain_abs + (d ^ 2u);
    } else if constexpr (LAY == LAY_BYTE) {
        ra.z = neg ? (uint32_t)add : 0u;
        ra.w = chain_abs + (uint32_t)ym;
    } else {
        ra.z = neg ? (uint32_t)add : 0u;
        ra.w = chain_abs;                                        // the owner site's byte
    }
    const uint32_t ksite = __umulhi(b << 3, (uint32_t)g.N);      // kick site :128
    rb.x = (LAY == LAY_NIB) ? (uint32_t)nib_fmt<CH>((int)ksite) : (uint32_t)g.SC * ksite;
    if constexpr (ALGO == ALGO_TAYLOR) {
        const bool dec = (b >> 29) & 1u;                         // :141
        uint32_t qi, qd;
        if (!g.lowt) {
            // T >= 1: kfix - 1 < 2^32 and tfix >= 2^32 > a.
            // (nb+1)*a < kfix  <=>  nb < floor((kfix-1)/a) =: q, capped at the storage limit 255.
            // q from a biased float quotient (an overestimate by < 1 whenever q < 257) and one exact
            // 64-bit check; a == 0 gives +inf -> saturated conversion -> cap, as it must.
            const uint32_t num = (uint32_t)(prm.kfix - 1);
            const uint32_t q0 = __float2uint_rz(__fdividef(numf, __uint2float_rn(a)));
            const unsigned long long pq = (unsigned long long)q0 * (unsigned long long)a;
            const uint32_t q = q0 - (uint32_t)(pq > (unsigned long long)num);
            qi = min(q, LAY == LAY_NIB ? 16u : 255u);            // (nibbles: 15 -> 16 is let through and caught at the end)
            qd = 1u;                                             // a < nb*tfix      <=>  nb >= 1
        } else {
            const unsigned long long num = prm.kfix - 1;
            const unsigned long long q1 = a ? num / (unsigned long long)a : 255ull;
            const unsigned long long q2 = (unsigned long long)a / (unsigned long long)prm.tfix + 1ull;   // nb > floor(a/tfix)
            constexpr unsigned long long CAPI = (LAY == LAY_NIB) ? 16ull : 255ull, CAPD = (LAY == LAY_NIB) ? 16ull : 256ull;
            qi = (uint32_t)(q1 < CAPI ? q1 : CAPI);
            qd = (uint32_t)(q2 < CAPD ? q2 : CAPD);
        }
        rb.y = (dec ? qd : qi) << sh4;
        rb.z = dec ? (0u - (1u << sh4)) : (1u << sh4);           // occupation change (in place); negative = flipped test
        rb.w = (dec ? 0xc0u : 0x80u) | (d << 2) | ((LAY == LAY_NIB) ? ((15u << sh4) << 8) : 0u);   // accept-log bits (accepted, decrement, direction) | nibble mask
    } else {
        rb.y = ((uint64_t)a < prm.hfix) ? 0u : 1u;               // accept iff nb >= thr (occupied: always)
        rb.z = 0xffffffffu;
        rb.w = 0x80u | (d << 2);
    }
}

// One worm step on a pre-decoded record.  On entry `head` already holds this step's kick when the
// worm is closed; on exit it holds the NEXT step's kick (kick_next) when the worm is closed then.
// alog: shared address of this step's log byte.
// LOGK < 0: the log bytes are stored one by one.  LOGK = 0..3 (full rounds, step j = LOGK mod 4): they are
// gathered in two registers and stored as one 32-bit word each after every fourth step -- a lone warp issues
// a shared-memory instruction every ~4 cycles, and the ones between the accept test and the next bond load
// are on the head -> head recurrence.
template <int ALGO, int LAY, int SCSH, bool HIST, bool SH, bool MEAS, int LOGK = -1>
__device__ __forceinline__ void lat_step(const LatGeo &g, unsigned long long hist_row, uint32_t hist_writer,
                                         LatState &c, const uint4 ra, const uint4 rb, const uint4 rc, const int kick_next,
                                         const uint32_t p_next, const uint32_t alog, const uint32_t dummy = 0u)
{
    constexpr bool DUAL = (LAY == LAY_DUAL);
    const int add = (int)ra.x, keep = (int)ra.y;
    const int kick = (int)rb.x;
    const uint32_t thr = rb.y, flip = rb.z, logbits = (LAY == LAY_NIB) ? (rb.w & 0xffu) : rb.w;
    const bool closed = c.closed != 0u;
    const int tail = closed ? kick : c.tail;                    // kick :128-129 (head has it already)
    c.tail = tail;
    const int head = c.head;
    uint32_t addr, addr2 = 0u;
    const int t = head + add;
    const int nh = (head & keep) | (t & ~keep);                 // neighbour table :88-95
    uint32_t nb;
    if constexpr (DUAL) {
        // bond_number :214-244: slot d of this site (address head + ra.z, formed and loaded by the previous
        // step as soon as its head was selected: the load is the long pole of the head -> head recurrence,
        // so it is issued right behind that step's stores instead of behind this step's address arithmetic)
        addr = c.addr;
        nb = c.nb;                                              // :139
        addr2 = (uint32_t)nh + ra.w;                            // ... and slot d^2 of the neighbour
    } else {
        const int u = head + (int)ra.z;
        const int own = (head & keep) | (u & ~keep);
        addr = (uint32_t)own + ra.w;
        nb = lds_u8(addr);                                      // :139  (nibble layout: the site's byte, both occupations)
    }
    // nibble layout: the occupation under test stays in place (masked), threshold and change come pre-shifted
    const uint32_t whole = nb;
    if constexpr (LAY == LAY_NIB) nb &= (rb.w >> 8);
    // Solo mode with histograms: G(r) bin of (head - tail) before the move; a closed worm is bin 0
    // (dh == 0).  The chain's 16-bit bins live in its own bank and only this lane touches them between
    // the walker's flushes, so the count is a plain load / add / store (a shared-memory atomic costs
    // the LSU ~2 cycles per lane); every lane stores (a slot that does not write adds 0).  The load
    // rides in the bond load's shadow, the add and the store follow the accept block.
    uint32_t hv = 0u, ha = 0u, dnew = 0u;
    if constexpr (HIST && SH) {
        ha = (uint32_t)hist_row + c.dh;
        if constexpr (MEAS) hv = lds_u16(ha);
        const uint32_t th = (c.dh | rc.z) + rc.x;
        dnew = (c.dh & rc.y) | (th & ~rc.y);
    }
    // closed iteration (worm_ising_2d.cpp:130-132): logged, measured by the accountant.  After the
    // load in program order: a store ahead of it would sit on the load's issue path (the compiler
    // cannot reorder shared-memory accesses that may alias).
    const uint32_t closed2 = c.closed << 1;                     // bit 1 of the step's log byte
    const uint32_t pc = (nh == tail) ? 1u : 0u;                 // closed after an accepted move?
    const int A = pc ? kick_next : nh;                          // head if accepted (kick folded in)
    const int B = closed ? kick_next : head;                    // head if rejected
    // thr + x * 0, written so that the compiler has to finish A, B and the second store address before
    // the first consumer of nb: the warp issues in order and stalls at that consumer, so everything
    // that does not depend on the load must sit in front of it (in the load's shadow), not behind
    const uint32_t thrx = thr + (uint32_t)(A ^ B ^ (int)addr2 ^ (int)dnew) * (uint32_t)g.zero;   // IMAD: the FMA pipe is the idle one
    const uint32_t nn = (ALGO == ALGO_TAYLOR) ? whole + flip : (nb ^ 1u);
    const uint32_t lg = (nb & (LAY == LAY_NIB ? 0x11u : 1u)) | logbits;
    // accept test (:151), head select, bond stores (:152-154) and the accept-log byte in one opaque
    // block: the compiler must not re-associate "closed ? kick : ..." across the select.
    int head_new;
    uint32_t acc_i;
    if constexpr (DUAL && LOGK < 0) {
        asm volatile("{\n\t.reg .pred pa, pd;\n\t.reg .b32 lb;\n\t"
                     "setp.lt.u32 pa, %4, %5;\n\t"
                     "setp.lt.s32 pd, %6, 0;\n\t"
                     "xor.pred pa, pa, pd;\n\t"
                     "selp.b32 %0, %7, %8, pa;\n\t"
                     "add.u32 %2, %0, %14;\n\t"
                     "selp.u32 %1, 1, 0, pa;\n\t"
                     "@pa st.shared.u8 [%9], %11;\n\t"
                     "@pa st.shared.u8 [%10], %11;\n\t"
                     "ld.shared.u8 %3, [%2];\n\t"
                     "selp.b32 lb, %13, 0, pa;\n\t"
                     "or.b32 lb, lb, %15;\n\t"
                     "st.shared.u8 [%12], lb;\n\t}"
                     : "=&r"(head_new), "=&r"(acc_i), "=&r"(c.addr), "=&r"(c.nb)
                     : "r"(nb), "r"(thrx), "r"(flip), "r"(A), "r"(B), "r"(addr), "r"(addr2), "r"(nn), "r"(alog), "r"(lg),
                       "r"(p_next), "r"(closed2)
                     : "memory");
    } else if constexpr (DUAL) {
        // Full rounds.  The two bond stores are unpredicated and take their ADDRESS by select: a guard
        // predicate is ready ~13 cycles after its ISETP, a select operand after 4, and the next bond load has
        // to follow the stores (the next step may walk straight back).  A rejected move writes the occupation
        // it would have stored to `dummy`, the padding byte of the chain's log row.
        uint32_t lgs = ((LOGK == 0) ? 0u : c.aword) + (closed2 << (8 * LOGK));
        asm volatile("{\n\t.reg .pred pa, pd;\n\t.reg .b32 nv, nw;\n\t"
                     "setp.lt.u32 pa, %5, %6;\n\t"
                     "setp.lt.s32 pd, %7, 0;\n\t"
                     "xor.pred pa, pa, pd;\n\t"
                     "selp.b32 %0, %8, %9, pa;\n\t"
                     "selp.b32 nv, %10, %15, pa;\n\t"
                     "selp.b32 nw, %11, %15, pa;\n\t"
                     "st.shared.u8 [nv], %12;\n\t"
                     "st.shared.u8 [nw], %12;\n\t"
                     "add.u32 %2, %0, %14;\n\t"
                     "ld.shared.u8 %3, [%2];\n\t"
                     "selp.u32 %1, 1, 0, pa;\n\t"
                     "@pa mad.lo.u32 %4, %13, %16, %4;\n\t}"        // accept-log byte into its word, under the guard
                     : "=&r"(head_new), "=&r"(acc_i), "=&r"(c.addr), "=&r"(c.nb), "+r"(lgs)
                     : "r"(nb), "r"(thrx), "r"(flip), "r"(A), "r"(B), "r"(addr), "r"(addr2), "r"(nn), "r"(lg), "r"(p_next), "r"(dummy),
                       "n"(1u << (8 * LOGK))
                     : "memory");
        c.aword = lgs;
    } else if constexpr (LOGK >= 0) {
        uint32_t lgs;
        asm volatile("{\n\t.reg .pred pa, pd;\n\t"
                     "setp.lt.u32 pa, %3, %4;\n\t"
                     "setp.lt.s32 pd, %5, 0;\n\t"
                     "xor.pred pa, pa, pd;\n\t"
                     "selp.b32 %0, %6, %7, pa;\n\t"
                     "selp.u32 %1, 1, 0, pa;\n\t"
                     "@pa st.shared.u8 [%8], %9;\n\t"
                     "selp.b32 %2, %10, 0, pa;\n\t}"
                     : "=&r"(head_new), "=&r"(acc_i), "=&r"(lgs)
                     : "r"(nb), "r"(thrx), "r"(flip), "r"(A), "r"(B), "r"(addr), "r"(nn), "r"(lg)
                     : "memory");
        c.aword = (LOGK == 0 ? 0u : c.aword) + ((lgs | closed2) << (8 * LOGK));
    } else {
        asm volatile("{\n\t.reg .pred pa, pd;\n\t.reg .b32 lb;\n\t"
                     "setp.lt.u32 pa, %2, %3;\n\t"
                     "setp.lt.s32 pd, %4, 0;\n\t"
                     "xor.pred pa, pa, pd;\n\t"
                     "selp.b32 %0, %5, %6, pa;\n\t"
                     "selp.u32 %1, 1, 0, pa;\n\t"
                     "@pa st.shared.u8 [%7], %8;\n\t"
                     "selp.b32 lb, %10, 0, pa;\n\t"
                     "or.b32 lb, lb, %11;\n\t"
                     "st.shared.u8 [%9], lb;\n\t}"
                     : "=&r"(head_new), "=&r"(acc_i)
                     : "r"(nb), "r"(thrx), "r"(flip), "r"(A), "r"(B), "r"(addr), "r"(nn), "r"(alog), "r"(lg), "r"(closed2)
                     : "memory");
    }
    c.head = head_new;
    c.closed = acc_i ? pc : c.closed;
    if constexpr (LOGK == 3) {
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(alog - 3u), "r"(c.aword) : "memory");
    }
    if constexpr (HIST && SH) {
        c.dh = acc_i ? dnew : c.dh;
        if constexpr (MEAS) sts_u16(ha, hv + hist_writer);
    }
}

// Replay of one round's log by the producer warp: lane (q, part) owns steps [part*SP, part*SP+SP) of
// chain slot q.  Occupation-sum / parity-count changes are prefix-summed across the parts, then
// every lane measures its closed steps (:130-132) with the exact Nb / n of that step.
// HIST: the accountants also keep G(r).  The accept-log byte carries the direction of the move (bits 2-3, from the
// step record), so the displacement head - tail can be replayed exactly like Nb: it changes on accepted moves
// only (a kick moves head and tail together; a worm that closes has walked its displacement back to 0 mod L).
// Every measured iteration counts at the displacement it starts with; a lane merges the iterations of its steps
// that share a bin and adds them with one `red` -- the walker does nothing for the histogram.
template <int ALGO, int CH, bool SH, bool HIST>
__device__ __forceinline__ void account_round(Account &k, uint32_t logbuf_abs, int q, int part, bool meas, int jb, int je, int L, int lgL)
{
    using Cfg = LatCfg<CH, SH>;
    constexpr int SP = Cfg::SP, PARTS = Cfg::PARTS, ACH = Cfg::ACH;
    static_assert(SP == 4 || SP == 2, "a lane's steps are one 32- or 16-bit word of the chain's log row");
    // this lane's SP steps are consecutive bytes of the chain's row: one load per log, one store to clear
    const uint32_t row = logbuf_abs + (uint32_t)(q * Cfg::LSTR + part * SP);
    uint32_t aw;
    if constexpr (SP == 4) {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(aw) : "r"(row) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(row), "r"(0u) : "memory");
    } else {
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(aw) : "r"(row) : "memory");
        asm volatile("st.shared.u16 [%0], %1;" :: "r"(row), "r"(0u) : "memory");
    }
    const uint32_t cw = aw & 0x02020202u;                          // closed-at-start flags of the lane's steps
    int dNb[SP], dn[SP];
    int sNb = 0, sn = 0;
#pragma unroll
    for (int i = 0; i < SP; ++i) {
        const uint32_t al = (aw >> (8 * i)) & 0xffu;
        const int accd = (int)(al >> 7);
        const int par = (al & 0x11u) ? -1 : 1;                       // old parity: bit 0 (bit 4 for a high nibble)
        dn[i] = accd * par;
        dNb[i] = (ALGO == ALGO_TAYLOR) ? accd * (1 - (int)((al >> 5) & 2u)) : dn[i];
        sNb += dNb[i]; sn += dn[i];
    }
    // inclusive scan over the parts of a chain (lanes aq, aq+ACH, aq+2ACH, ...)
    int pNb = sNb, pn = sn;
#pragma unroll
    for (int o = ACH; o < 32; o <<= 1) {
        const int uNb = __shfl_up_sync(0xffffffffu, pNb, o);
        const int un = __shfl_up_sync(0xffffffffu, pn, o);
        if (part * ACH >= o) { pNb += uNb; pn += un; }
    }
    int Nb = k.Nb + pNb - sNb, npar = k.npar + pn - sn;          // state at this lane's first step
    const int last = (PARTS - 1) * ACH + (q % ACH);
    k.Nb += __shfl_sync(0xffffffffu, pNb, last);
    k.npar += __shfl_sync(0xffffffffu, pn, last);
    if (meas && cw) {
#pragma unroll
        for (int i = 0; i < SP; ++i) {
            if ((cw >> (8 * i)) & 0xffu) {                           // closed iteration :130-132
                k.Z += 1;
                k.SNb += (uint32_t)Nb;
                k.SNb2 += (uint64_t)(uint32_t)Nb * (uint32_t)Nb;
                k.Sn += (uint32_t)npar;
                k.Sn2 += (uint64_t)(uint32_t)npar * (uint32_t)npar;
            }
            Nb += dNb[i]; npar += dn[i];
        }
    }
    if constexpr (HIST) {
        // displacement changes of this lane's steps, packed (dx in the low half, dy in the high half: sums of +-1
        // over a round cannot reach the other half)
        int pd[SP];
        int spd = 0;
#pragma unroll
        for (int i = 0; i < SP; ++i) {
            const uint32_t al = (aw >> (8 * i)) & 0xffu;
            const int dir = (int)((al >> 2) & 3u);
            const int unit = (dir & 1) ? 65536 : 1;
            pd[i] = (al >> 7) ? ((dir & 2) ? -unit : unit) : 0;
            spd += pd[i];
        }
        int ppd = spd;
#pragma unroll
        for (int o = ACH; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ppd, o);
            if (part * ACH >= o) ppd += u;
        }
        int off = ppd - spd;                                      // displacement change before this lane's first step
        const int tot = __shfl_sync(0xffffffffu, ppd, last);
        if (meas && k.hrow) {
            // Branch-free: a run starts at a valid step that is the lane's first, or follows an accepted move or an
            // invalid step; its length is the number of valid steps up to the next accepted move.  One predicated
            // `red` per possible run start.
            const int mask = L - 1;
            bool v[SP], ac[SP];
            int len[SP + 1];
            len[SP] = 0;
#pragma unroll
            for (int i = 0; i < SP; ++i) {
                const int j = part * SP + i;
                v[i] = (j >= jb) && (j < je);
                ac[i] = ((aw >> (8 * i + 7)) & 1u) != 0u;
            }
#pragma unroll
            for (int i = SP - 1; i >= 0; --i) len[i] = v[i] ? 1 + (ac[i] ? 0 : len[i + 1]) : 0;
#pragma unroll
            for (int i = 0; i < SP; ++i) {
                const bool start = v[i] && (i == 0 || ac[i - 1] || !v[i - 1]);
                if (start) {
                    const int x = (k.dx + (int)(short)(off & 0xffff)) & mask;
                    const int y = (k.dy + ((off + 0x8000) >> 16)) & mask;
                    atomicAdd(k.hrow + ((y << lgL) | x), (unsigned long long)len[i]);      // hist[dy * L + dx], before the kick
                }
                off += pd[i];
            }
        }
        k.dx = (k.dx + (int)(short)(tot & 0xffff)) & (L - 1);
        k.dy = (k.dy + ((tot + 0x8000) >> 16)) & (L - 1);
    }
}

// The round schedule every role follows: rounds start on a Philox pair, never straddle a segment
// (MEAS change / dump multiple), and hold at most R steps.  Consecutive FULL rounds (the steady
// state) are handed out as one batch so that the per-round bookkeeping is a counter.
struct Batch {
    uint64_t s0;     // base step of the first round's record buffer (even)
    int jb, je;      // the first round walks steps [s0 + jb, s0 + je); all later rounds of the batch are full
    uint32_t n;      // rounds in the batch (n > 1 only for full rounds)
    uint64_t s_next; // first step after the batch
};
template <int R>
__device__ __forceinline__ Batch make_batch(uint64_t s, uint64_t seg_end)
{
    Batch b;
    b.s0 = s & ~1ull;
    b.jb = (int)(s - b.s0);
    const uint64_t left = seg_end - b.s0;
    const int rem = (int)(left % (uint64_t)R);
    if (b.jb == 0 && left > (uint64_t)R && rem != 0 && (rem & 1) == 0) {
        // The short round of a segment comes FIRST (when it is even: rounds start on a Philox pair).  With the dump
        // rule every 50-step interval has one, and two record buffers keep the roles in lockstep: behind the short
        // round the walker waits for a whole round to be built.  In front of the interval that wait overlaps the
        // walker's pause at the dump multiple.
        b.n = 1u;
        b.je = rem;
        b.s_next = b.s0 + (uint64_t)rem;
    } else if (b.jb == 0 && left >= (uint64_t)R) {
        const uint64_t nf = left / (uint64_t)R;
        b.n = nf > 0x40000000ull ? 0x40000000u : (uint32_t)nf;
        b.je = R;
        b.s_next = b.s0 + (uint64_t)b.n * (uint64_t)R;
    } else {
        b.n = 1u;
        b.je = left < (uint64_t)R ? (int)left : R;
        b.s_next = b.s0 + (uint64_t)b.je;
    }
    return b;
}

// ---- bond storage <-> the caller's compact uint8 [2N] array (bond 2*site = +x, 2*site+1 = +y)
// dual: word of site s of chain slot q (within its walker) at region + (s * CH + q) * 4
template <int CH>
__device__ __forceinline__ uint32_t dual_word(const uint8_t *gb, int s, int L, int N)
{
    const int x = s & (L - 1);
    const int left = (x == 0) ? s + L - 1 : s - 1;
    const int down = (s < L) ? s + N - L : s - L;
    const uint32_t xy = *reinterpret_cast<const uint16_t *>(gb + 2 * s);        // +x | +y << 8
    return xy | ((uint32_t)gb[2 * left] << 16) | ((uint32_t)gb[2 * down + 1] << 24);
}

// Dump copy (worm_ising_2d.cpp:120-124): the walker's live lanes copy, chain by chain, the bonds of the chains in `dm`
// to their next dump slot.  Cold code, kept out of line.
template <int CH, int LAY>
__device__ __noinline__ void lat_copy_dumps(unsigned dm, int lane, long long cl, int nd, uint8_t *dumps, int max_dumps,
                                            int nb2, int N, uint32_t region_abs)
{
    constexpr unsigned wmask = (CH == 32) ? 0xffffffffu : ((1u << CH) - 1u);
    while (dm) {
        const int src_lane = __ffs(dm) - 1;
        dm &= dm - 1;
        const long long chain = __shfl_sync(wmask, cl, src_lane);
        const int slot = __shfl_sync(wmask, nd, src_lane);
        uint8_t *dstb = dumps + (chain * max_dumps + slot) * nb2;
        if constexpr (LAY == LAY_DUAL) {
            uint16_t *dst = reinterpret_cast<uint16_t *>(dstb);
            for (int sI = lane; sI < N; sI += CH)
                dst[sI] = (uint16_t)lds_u32(region_abs + (uint32_t)((sI * CH + src_lane) * 4));
        } else if constexpr (LAY == LAY_NIB) {
            uint2 *dst = reinterpret_cast<uint2 *>(dstb);       // four sites = one word of the chain's column = 8 bond bytes
            for (int wI = lane; wI < N / 4; wI += CH) {
                const uint32_t v = lds_u32(region_abs + (uint32_t)(wI * 4 * CH + src_lane * 4));
                const uint32_t lo = v & 0x0f0f0f0fu, hi = (v >> 4) & 0x0f0f0f0fu;      // +x / +y occupations of the four sites
                dst[wI] = make_uint2(__byte_perm(lo, hi, 0x5140), __byte_perm(lo, hi, 0x7362));
            }
        } else {
            uint4 *dst = reinterpret_cast<uint4 *>(dstb);
            for (int i = lane; i < nb2 / 16; i += CH) dst[i] = lds_v4(region_abs + (uint32_t)(i * 16));
        }
    }
}

// HMODE: G(r) histograms -- 0 none; 1 the solo walker counts in 16-bit per-chain bins in shared memory (needs
// 2 N bytes per chain: L <= 32); 2 the accountants replay the displacement from the logged directions and add
// runs of equal bins with `red.global` (any lattice, no walker work, but bounded by the L2's atomic rate).
template <int ALGO, int CH, int LAY, int HMODE>
__global__ void __launch_bounds__(512, 1)
worm_lat_kernel(const PhiloxArgs a, const LatGeo g)
{
    constexpr bool HIST = (HMODE == 2);                 // accountant histograms
    constexpr bool DUAL = (LAY == LAY_DUAL);
    constexpr bool NIB = (LAY == LAY_NIB);
    constexpr bool SH = (HMODE == 1);                   // per-chain G(r) bins in shared memory, counted by the walker
    static_assert(!SH || CH == 32, "walker-side histograms are a solo-mode feature");
    using Cfg = LatCfg<CH, SH>;
    constexpr int R = Cfg::R, NP = Cfg::NP;
    constexpr int SC = DUAL ? 4 * CH : NIB ? CH : 2;    // bytes per unit of the scaled site index
    constexpr int SCSH = (SC == 1) ? 0 : (SC == 2) ? 1 : (SC == 4) ? 2 : (SC == 8) ? 3 : (SC == 16) ? 4 : (SC == 32) ? 5 : 7;
    static_assert(LAY != LAY_BYTE || CH == 1, "the byte-per-bond layout is one chain per walker");
    static_assert(!NIB || ALGO == ALGO_TAYLOR, "nibble layout: Taylor chain only");
    extern __shared__ __align__(16) uint8_t smem8[];
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int nw = g.nw;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // Roles by warp.  CH <= 8: role r = warp / nw of group w = warp % nw -- r < NP producer of Philox
    // block r, r == NP accountant, r == NP+1 walker; the warps of a group share an SM sub-partition
    // (nw = 4) and the walker is its highest warp id (the arbiter prefers it).
    // CH == 32 ("solo"): one group per block; the walker (warp 0) has sub-partition 0 to itself --
    // warps 4, 8, 12 stay idle -- and the twelve helpers (8 producers, 4 accountants of 8 chains each)
    // live on the other three sub-partitions.
    enum { PRODUCER, ACCOUNTANT, WALKER, IDLE };
    constexpr int NA = Cfg::NA, ACH = Cfg::ACH;
    int kind, u = 0, w = 0;
    if constexpr (CH == 32) {
        if (warp == 0) kind = WALKER;
        else if ((warp & 3) == 0) kind = IDLE;
        else {
            const int h = warp - 1 - (warp >> 2);            // 0..11 for warps 1,2,3,5,6,7,9,10,11,13,14,15
            kind = h < NP ? PRODUCER : (h - NP < NA ? ACCOUNTANT : IDLE);
            u = h < NP ? h : h - NP;                         // producer: Philox block; accountant: which 8 chains
        }
    } else {
        const int role = warp / nw;
        w = warp - role * nw;
        kind = role < NP ? PRODUCER : (role < NP + NA ? ACCOUNTANT : WALKER);
        u = role < NP ? role : role - NP;
    }
    const bool is_walker = kind == WALKER;
    const int bch = nw * CH;                             // chains per block
    const int64_t chain0 = (int64_t)blockIdx.x * bch;    // first chain of the block
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem8);
    const uint32_t rec_w = sbase + (uint32_t)g.rec_off + (uint32_t)(w * Cfg::RECW);
    const uint32_t log_w = sbase + (uint32_t)g.log_off + (uint32_t)(w * Cfg::LOGW);
    const uint32_t region_abs = sbase + (uint32_t)(w * g.region);
    const int bar_id = 1 + w;

    // ---- stage the block's bonds, clear the step logs (all warps)
    if constexpr (DUAL) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(smem8);
        const int total = bch * g.N;                     // words; index = (walker * N + s) * CH + q
        for (int i = tid; i < total; i += nthreads) {
            const int q = i % CH;
            const int ws = i / CH;
            const int wk = ws / g.N, s = ws - wk * g.N;
            const int64_t cg = chain0 + wk * CH + q;
            dst[i] = (cg < a.n_chains) ? dual_word<CH>(a.bonds + cg * (int64_t)g.nb2, s, g.L, g.N) : 0u;
        }
    } else if constexpr (NIB) {
        // byte walker * region + nib_fmt(s) + 4 q = occupations of site s of chain slot q: +x | +y << 4
        uint8_t *bad = smem8 + g.bad_off;
        for (int i = tid; i < bch; i += nthreads) bad[i] = 0;
        __syncthreads();
        const int total = bch * g.N;
        for (int i = tid; i < total; i += nthreads) {
            const int q = i % CH;
            const int ws = i / CH;
            const int wk = ws / g.N, s = ws - wk * g.N;
            const int64_t cg = chain0 + wk * CH + q;
            uint32_t xy = 0u;
            if (cg < a.n_chains) xy = *reinterpret_cast<const uint16_t *>(a.bonds + cg * (int64_t)g.nb2 + 2 * s);
            if (xy & 0xf0f0u) bad[wk * CH + q] = 1;              // an occupation above 15 does not fit
            smem8[wk * g.region + nib_fmt<CH>(s) + 4 * q] = (uint8_t)((xy & 15u) | ((xy >> 4) & 0xf0u));
        }
    } else {
        const int vec_per_chain = g.nb2 / 16;
        const int total = bch * vec_per_chain;
        uint4 *dst = reinterpret_cast<uint4 *>(smem8);
        for (int i = tid; i < total; i += nthreads) {
            const int cslot = i / vec_per_chain;
            const int64_t cg = chain0 + cslot;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cg < a.n_chains)
                v = reinterpret_cast<const uint4 *>(a.bonds + cg * (int64_t)g.nb2)[i - cslot * vec_per_chain];
            dst[i] = v;
        }
    }
    {
        uint32_t *lg = reinterpret_cast<uint32_t *>(smem8 + g.log_off);
        for (int i = tid; i < nw * Cfg::LOGW / 4; i += nthreads) lg[i] = 0u;
        // the record rings too: the last step of a launch forms (and loads) the address of a step that no
        // producer has built -- 0 keeps that address inside the block's bond storage
        uint32_t *rr = reinterpret_cast<uint32_t *>(smem8 + g.rec_off);
        for (int i = tid; i < nw * Cfg::RECW / 4; i += nthreads) rr[i] = 0u;
        if constexpr (SH) {
            if (tid < 4) *reinterpret_cast<uint4 *>(smem8 + g.htab_off + tid * 16) = hist_entry(g, (uint32_t)tid);
            uint32_t *hb = reinterpret_cast<uint32_t *>(smem8 + g.hist_off);
            for (int i = tid; i < CH * g.N / 2; i += nthreads) hb[i] = 0u;
        }
    }
    __syncthreads();

    if (kind == PRODUCER) {
        // ================================ producer warp ================================
        // Philox block b = u*32 + lane of a round (u = role: the group's two producers take one
        // block each) belongs to chain slot q = b % CH, pair p = b / CH: consecutive lanes store
        // consecutive 16-byte records (conflict-free 128-bit stores).
        const int b = u * 32 + lane;
        const int q = b % CH;
        const int p_of = b / CH;
        const int64_t cg = chain0 + w * CH + q;
        const int64_t cl = cg < a.n_chains ? cg : 0;
        const PhiloxChain prm = a.prm[cl];
        const uint64_t chain_id = a.chain_id0 + (uint64_t)cl;
        const uint32_t c2 = (uint32_t)chain_id, c3 = (uint32_t)(chain_id >> 32);
        const uint32_t chain_abs = region_abs + ((DUAL || NIB) ? (uint32_t)q * 4u : 0u);
        const uint32_t rec_q = rec_w + (uint32_t)q * 16u + (uint32_t)(2 * p_of) * (CH * 16);
        const PhiloxKeys keys = philox_keys((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32));
        const float numf = __fmul_ru(__uint2float_ru((uint32_t)(prm.kfix - 1)), 1.0f + 1.0f / 262144.0f);
        uint32_t buf = 0u;
        // steps of a round this lane's record(s) serve: a short round (segment edge) is built by the lanes it needs
        const int j_lo = 2 * p_of, j_hi = 2 * p_of + 1;
        // the records of one Philox pair per lane into the current buffer
        auto build = [&](uint64_t pair) {
#if WORM_LAT_FAKE   // measurement aid only (wrong random numbers): how much of the walker's step waits for the producers
            uint4 r = make_uint4((uint32_t)pair * 2654435761u, (uint32_t)pair * 40503u + c2, c2 * 2246822519u + (uint32_t)pair, c3 + (uint32_t)pair * 3266489917u);
            r.x ^= r.y >> 7; r.y ^= r.x << 9; r.z ^= r.w >> 5; r.w ^= r.z << 11;
#else
            const uint4 r = philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), c2, c3, keys);
#endif
            const uint32_t p0 = rec_q + buf;
            uint4 ra0, rb0, ra1, rb1;
            make_record<ALGO, LAY, CH>(g, prm, numf, r.x, r.y, chain_abs, ra0, rb0);
            make_record<ALGO, LAY, CH>(g, prm, numf, r.z, r.w, chain_abs, ra1, rb1);
            sts_v4(p0, ra0); sts_v4(p0 + Cfg::REC_B, rb0);
            sts_v4(p0 + CH * 16, ra1); sts_v4(p0 + CH * 16 + Cfg::REC_B, rb1);
        };
        const bool dump_even = a.dump_every != 0ull && (a.dump_every & 1ull) == 0ull;
        const int dump_full = (int)(a.dump_every / (uint64_t)R), dump_tail = (int)(a.dump_every % (uint64_t)R);
        Segments segs(a);
        uint64_t s = a.step_begin;
        while (s < a.step_end) {
            if (dump_even && s == segs.next_dump && a.step_end - s >= a.dump_every) {
                // whole dump intervals in the steady state (the schedule of make_batch: an even short round, then the
                // full rounds) as one tight loop: the segment bookkeeping below, cold code once per 50 steps, made the
                // walker wait ~500 cycles for the first round of every interval
                uint64_t pair = (s >> 1) + (uint64_t)p_of;
                do {
                    const int nr = dump_full + (dump_tail && dump_full ? 1 : 0) + (dump_full ? 0 : 1);
                    for (int i = 0; i < nr; ++i) {
                        const int je = (i == 0 && (dump_tail || !dump_full)) ? (dump_tail ? dump_tail : R) : R;
                        if (j_lo < je) build(pair);
                        group_barrier<Cfg::GW>(bar_id);
                        buf ^= (uint32_t)Cfg::BUF;
                        pair += (uint64_t)(je >> 1);
                    }
                    s += a.dump_every;
                    segs.next_dump += a.dump_every;
                } while (a.step_end - s >= a.dump_every);
                continue;
            }
            bool meas, dump_first;
            const uint64_t seg_end = segs.next(a, s, meas, dump_first);
            while (s < seg_end) {
                const Batch bt = make_batch<R>(s, seg_end);
                uint64_t pair = (bt.s0 >> 1) + (uint64_t)p_of;
                for (uint32_t kk = 0; kk < bt.n; ++kk, pair += R / 2) {
                    if (j_hi >= bt.jb && j_lo < bt.je) build(pair);
                    group_barrier<Cfg::GW>(bar_id);
                    buf ^= (uint32_t)Cfg::BUF;
                }
                s = bt.s_next;
            }
        }
        group_barrier<Cfg::GW>(bar_id);                            // matches the walker's barrier before its last steps
        group_barrier<Cfg::GW>(bar_id);                            // the walker has finished its last step
    } else if (kind == ACCOUNTANT) {
        // ================================ accountant warp ================================
        // accountant identity of this lane: chain slot aq, part ap of that chain's steps
        const int aq = u * ACH + lane % ACH, ap = lane / ACH;
        const int64_t acg = chain0 + w * CH + aq;
        bool alive = acg < a.n_chains;
        const int64_t acl = alive ? acg : 0;
        if constexpr (NIB) {
            // a chain whose input does not fit 4 bits per bond is reported and left untouched
            if (smem8[g.bad_off + w * CH + aq]) {
                if (alive && ap == 0) a.acc[acl * ACC_STRIDE + ACC_STATUS] |= 1;
                alive = false;
            }
        }
        Account k;
        k.Z = k.SNb = k.SNb2 = k.Sn = k.Sn2 = 0;
        int64_t *ac = a.acc + acl * ACC_STRIDE;
        uint64_t old[5] = {0, 0, 0, 0, 0};
        if (ap == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) old[i] = (uint64_t)ac[i];
            k.Z = old[0]; k.SNb = old[1]; k.SNb2 = old[2]; k.Sn = old[3]; k.Sn2 = old[4];
        }
        k.nd = (a.n_dumps && alive) ? a.n_dumps[acl] : 0;
        k.dx = k.dy = 0;
        k.hrow = nullptr;
        int lgL = 0;
        while ((1 << lgL) < g.L) ++lgL;
        if constexpr (HIST) {
            const int hs = a.head[acl], ts = a.tail[acl];
            k.dx = (hs - ts) & (g.L - 1);
            k.dy = ((hs >> lgL) - (ts >> lgL)) & (g.L - 1);
            if (alive && a.hist && a.group_index)
                k.hrow = reinterpret_cast<unsigned long long *>(a.hist + (int64_t)a.group_index[acl] * g.N);
        }
        {   // occupation sum / odd-bond count of the staged configuration: the chain's parts share the sites
            int Nb = 0, npar = 0;
            if constexpr (DUAL) {
                for (int s = ap; s < g.N; s += Cfg::PARTS) {
                    const uint32_t v = lds_u32(region_abs + (uint32_t)((s * CH + aq) * 4)) & 0xffffu;
                    Nb += (int)__vsadu4(v, 0);
                    npar += __popc(v & 0x0101u);
                }
            } else if constexpr (NIB) {
                for (int s = ap; s < g.N; s += Cfg::PARTS) {
                    const uint32_t v = lds_u8(region_abs + (uint32_t)(nib_fmt<CH>(s) + 4 * aq));
                    Nb += (int)((v & 15u) + (v >> 4));
                    npar += (int)((v & 1u) + ((v >> 4) & 1u));
                }
            } else {
                for (int i = ap; i < g.nb2 / 16; i += Cfg::PARTS) {
                    const uint4 v = lds_v4(region_abs + (uint32_t)(i * 16));
                    Nb += (int)(__vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0));
                    npar += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                            __popc(v.w & 0x01010101u);
                }
            }
#pragma unroll
            for (int o = ACH; o < 32; o <<= 1) {
                Nb += __shfl_xor_sync(0xffffffffu, Nb, o);
                npar += __shfl_xor_sync(0xffffffffu, npar, o);
            }
            k.Nb = Nb; k.npar = npar;
        }
        // replay of a finished round: dump event first (values before the round's first step), then the log
        struct Pending { uint64_t s; int jb, je; bool meas, dump_first, valid; };
        Pending pend[2] = {{0, 0, 0, false, false, false}, {0, 0, 0, false, false, false}};
        uint32_t lbuf_acc = 0u;                          // log buffer of the oldest pending round
        auto account = [&](const Pending &pr) {
            const uint32_t lb = log_w + lbuf_acc;
            if (pr.dump_first) {                                     // worm_ising_2d.cpp:113-119
                const uint32_t was_closed = lds_u8(lb + (uint32_t)(aq * Cfg::LSTR + pr.jb)) & 2u;
                uint64_t Z = k.Z, SNb = k.SNb;
#pragma unroll
                for (int o = ACH; o < 32; o <<= 1) {
                    Z += __shfl_xor_sync(0xffffffffu, Z, o);
                    SNb += __shfl_xor_sync(0xffffffffu, SNb, o);
                }
                if (was_closed) {
                    if (ap == 0 && alive && k.nd < a.max_dumps && a.events) {
                        int64_t *e = a.events + ((int64_t)acl * a.max_dumps + k.nd) * 3;
                        e[0] = (int64_t)pr.s; e[1] = (int64_t)Z; e[2] = (int64_t)SNb;
                    }
                    ++k.nd;
                }
            }
            account_round<ALGO, CH, SH, HIST>(k, lb, aq, ap, pr.meas, pr.jb, pr.je, g.L, lgL);
            lbuf_acc = (lbuf_acc == 2u * Cfg::LOGB) ? 0u : lbuf_acc + (uint32_t)Cfg::LOGB;
        };

        const bool dump_even = a.dump_every != 0ull && (a.dump_every & 1ull) == 0ull;
        const int dump_full = (int)(a.dump_every / (uint64_t)R), dump_tail = (int)(a.dump_every % (uint64_t)R);
        Segments segs(a);
        uint64_t s = a.step_begin;
        while (s < a.step_end) {
            if (dump_even && s == segs.next_dump && a.step_end - s >= a.dump_every) {
                // whole dump intervals in the steady state, like the producers
                do {
                    const int nr = dump_full + (dump_tail && dump_full ? 1 : 0) + (dump_full ? 0 : 1);
                    uint64_t sr = s;
                    for (int i = 0; i < nr; ++i) {
                        const int je = (i == 0 && (dump_tail || !dump_full)) ? (dump_tail ? dump_tail : R) : R;
                        group_barrier<Cfg::GW>(bar_id);
                        if (pend[0].valid) account(pend[0]);
                        pend[0] = pend[1];
                        pend[1] = Pending{sr, 0, je, true, i == 0, true};
                        sr += (uint64_t)je;
                    }
                    s += a.dump_every;
                    segs.next_dump += a.dump_every;
                } while (a.step_end - s >= a.dump_every);
                continue;
            }
            bool meas, dump_first;
            const uint64_t seg_end = segs.next(a, s, meas, dump_first);
            bool seg_first = true;
            while (s < seg_end) {
                const Batch bt = make_batch<R>(s, seg_end);
                uint64_t sr = s;
                for (uint32_t kk = 0; kk < bt.n; ++kk) {
                    group_barrier<Cfg::GW>(bar_id);
                    // the walker is now inside the previous round: the one before that is complete
                    if (pend[0].valid) account(pend[0]);
                    pend[0] = pend[1];
                    pend[1] = Pending{sr, kk ? 0 : bt.jb, kk ? R : bt.je, meas, dump_first && seg_first, true};
                    seg_first = false;
                    sr = bt.s0 + (uint64_t)(kk + 1) * (uint64_t)R;
                }
                s = bt.s_next;
            }
        }
        group_barrier<Cfg::GW>(bar_id);                            // matches the walker's barrier before its last steps
        group_barrier<Cfg::GW>(bar_id);                            // the walker has finished its last step
        if (pend[0].valid) account(pend[0]);
        if (pend[1].valid) account(pend[1]);
        if constexpr (NIB) {
            // an occupation that reached 16 carried out of its nibble: the sum over the chain's bytes has lost 16
            // against the replayed sum of occupations
            int sum = 0;
            for (int s = ap; s < g.N; s += Cfg::PARTS) {
                const uint32_t v = lds_u8(region_abs + (uint32_t)(nib_fmt<CH>(s) + 4 * aq));
                sum += (int)((v & 15u) + (v >> 4));
            }
#pragma unroll
            for (int o = ACH; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (ap == 0 && alive && sum != k.Nb) a.acc[acl * ACC_STRIDE + ACC_STATUS] |= 2;
        }
        // reduce the parts' shares and write the chain's record
#pragma unroll
        for (int o = ACH; o < 32; o <<= 1) {
            k.Z += __shfl_xor_sync(0xffffffffu, k.Z, o);
            k.SNb += __shfl_xor_sync(0xffffffffu, k.SNb, o);
            k.SNb2 += __shfl_xor_sync(0xffffffffu, k.SNb2, o);
            k.Sn += __shfl_xor_sync(0xffffffffu, k.Sn, o);
            k.Sn2 += __shfl_xor_sync(0xffffffffu, k.Sn2, o);
        }
        if (ap == 0 && alive) {
            const int64_t iters = measured_iters(a);
            add_group_acc(a, acl, ac, k.Z, k.SNb, k.SNb2, k.Sn, k.Sn2, iters);
            ac[0] = (int64_t)k.Z; ac[1] = (int64_t)k.SNb; ac[2] = (int64_t)k.SNb2;
            ac[3] = (int64_t)k.Sn; ac[4] = (int64_t)k.Sn2;
            ac[5] += iters;
        }
    } else if (kind == WALKER) {
        // ================================= walker warp =================================
        // One lane per chain; the other lanes of the warp retire here (barriers count warps, and a
        // shared-memory access by CH lanes costs CH/8 wavefronts instead of 4).
        if (lane >= CH) return;
        constexpr unsigned wmask = (CH == 32) ? 0xffffffffu : ((1u << CH) - 1u);
        const int q = lane;                  // chain slot within the warp
        const int64_t cidx = chain0 + w * CH + q;
        bool live = cidx < a.n_chains;
        if constexpr (NIB) live = live && smem8[g.bad_off + w * CH + q] == 0;   // a chain that does not fit is left untouched
        const int64_t cl = live ? cidx : 0;  // dead slots shadow chain 0 in their own shared-memory slot, never write
        uint32_t rec_q = rec_w + (uint32_t)q * 16u;
        uint32_t log_q = log_w + (uint32_t)(q * Cfg::LSTR);
        asm volatile("" : "+r"(rec_q), "+r"(log_q));                     // keep them in registers
        unsigned long long hist_row = 0ull;
        unsigned long long *hist_out = nullptr;         // SH: this chain's row of the int64 histogram in HBM
        uint32_t hist_writer = 0u;
        if constexpr (SH) {
            // SH: every slot, live or not, has its own bank of bins in shared memory (adds of 0 go there too):
            // bins 2i, 2i+1 of slot q are the halves of word i * 32 + q
            if constexpr (SH) hist_row = (unsigned long long)(sbase + (uint32_t)g.hist_off + (uint32_t)(q * 4));
            if (live && a.hist && a.group_index) {
                unsigned long long *row = reinterpret_cast<unsigned long long *>(a.hist + (int64_t)a.group_index[cl] * g.N);
                if constexpr (SH) hist_out = row;
                else hist_row = (unsigned long long)row;
                hist_writer = 1u;
            }
        }
        // SH: the walker sweeps its chains' 16-bit bins into the int64 histogram every Cfg::FLUSH rounds
        // and at the end (lane q owns bank q: no conflicts, no atomics in shared memory)
        auto flush_hist = [&]() {
            if constexpr (SH) {
                const uint32_t base = (uint32_t)hist_row;
                for (int i0 = 0; i0 < g.N / 2; i0 += 4) {
                    uint32_t v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = lds_u32(base + (uint32_t)((i0 + e) * 128));
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (v[e]) {
                            asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + (uint32_t)((i0 + e) * 128)), "r"(0u) : "memory");
                            if (hist_out) {
                                if (v[e] & 0xffffu) atomicAdd(hist_out + 2 * (i0 + e), (unsigned long long)(v[e] & 0xffffu));
                                if (v[e] >> 16) atomicAdd(hist_out + 2 * (i0 + e) + 1, (unsigned long long)(v[e] >> 16));
                            }
                        }
                    }
                }
            }
        };
        int flush_ctr = 0;

        LatState c;
        c.head = NIB ? nib_fmt<CH>(a.head[cl]) : SC * a.head[cl];
        c.tail = NIB ? nib_fmt<CH>(a.tail[cl]) : SC * a.tail[cl];
        c.closed = (c.head == c.tail) ? 1u : 0u;
        c.dh = 0u;
        c.addr = sbase; c.nb = 0u;
        c.aword = 0u;
        if constexpr (SH) {
            const int hs = a.head[cl], ts = a.tail[cl];
            const int dx = (hs - ts) & (g.L - 1);
            const int dy = (hs / g.L - ts / g.L) & (g.L - 1);
            const int sd = dy * g.L + dx;
            c.dh = (uint32_t)((sd >> 1) * 128 + (sd & 1) * 2);
        }
        int32_t nd = (a.n_dumps && live) ? a.n_dumps[cl] : 0;

#if WORM_LAT_PROF
        long long pf_seg = 0, pf_dump = 0, pf_full = 0, pf_tail = 0, pf_edge = 0, pf_n = 0, pf_t = clock64(), pf_t0 = pf_t;
#define PF(x) { const long long t_ = clock64(); x += t_ - pf_t; pf_t = t_; }
#else
#define PF(x)
#endif
        auto rec_addr = [&](uint32_t buf, int j) { return rec_q + buf + (uint32_t)(j * CH * 16); };
        // A dump multiple (worm_ising_2d.cpp:120-124): the chains that are closed in front of this step are copied
        // to their next dump slot (the event record {step, Z, sum Nb} is written by the accountant)
        auto do_pause = [&]() {
            const bool want = live && c.closed;
            const unsigned dm = __ballot_sync(wmask, want && nd < a.max_dumps && a.dumps != nullptr);
            if (dm) lat_copy_dumps<CH, LAY>(dm, lane, (long long)cl, nd, a.dumps, a.max_dumps, g.nb2, g.N, region_abs);
            if (want) ++nd;
        };

        // A full round (steps 0..R-1 of the buffer at `cur`), software-pipelined two steps deep: on
        // entry (p0a,p0b) / (p1a,p1b) hold the records of steps 0 / 1; step j fetches the record of
        // step j+2, which for the last two steps lives in the other buffer (complete once the pair
        // barrier has been passed).  On exit they hold the records of steps 0 / 1 of the next round.
        uint4 p0a, p0b, p1a, p1b;
        uint4 p0c = make_uint4(0, 0, 0, 0);          // SH only: displacement update of the record in (p0a, p0b)
        const uint32_t htab = sbase + (uint32_t)g.htab_off;
        // (looked up one step ahead, from the direction bits of the next step's record -- already in registers)
        auto hist_lookup = [&](uint32_t rbw) { return lds_v4(htab + ((rbw & 0xcu) << 2)); };
        auto full_round = [&](auto meas_tag, uint32_t cur, uint32_t nxt, uint32_t lg) {
            constexpr bool MEAS = decltype(meas_tag)::value;
            static_assert(R % 4 == 0, "log words hold four steps");
            uint32_t pad = lg + (uint32_t)R;         // padding byte of the chain's log row: where a rejected move's stores go
            asm volatile("" : "+r"(pad));            // (a register, not base + immediate: one address for all steps)
            // U steps are unrolled; a round longer than that (one chain per walker: 64 steps) loops over chunks of U --
            // 64 unrolled steps are 30 KB of code, which a lone warp streams from L2 on every round (~30 of its 100
            // cycles per step); 16 stay in the 32 KB L1.5.  For R == U the chunk loop and its tests fold away.
            constexpr int U = 16, NCH = R / U;
            static_assert(R % U == 0, "rounds are whole chunks");
#pragma unroll 1
            for (int h = 0; h < NCH; ++h) {
                const bool last = (NCH == 1) || (h == NCH - 1);
                const uint32_t rb0 = cur + (uint32_t)(h * U * CH * 16);          // this chunk's records
                const uint32_t rb1 = last ? nxt : rb0 + (uint32_t)(U * CH * 16);  // ... and the next chunk's
                const uint32_t lgh = lg + (uint32_t)(h * U);
                unroll_steps(std::make_integer_sequence<int, U>{}, [&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if (j == U - 2) {
                        if (last) group_barrier<Cfg::GW>(bar_id);
                    }
                    const uint32_t nrec = (j < U - 2) ? rec_addr(rb0, j + 2) : rec_addr(rb1, j + 2 - U);
                    const uint4 fa = lds_v4(nrec), fb = lds_v4(nrec + Cfg::REC_B);
                    uint4 fc = make_uint4(0, 0, 0, 0);
                    if constexpr (SH) fc = hist_lookup(p1b.w);
                    lat_step<ALGO, LAY, SCSH, SH, SH, MEAS, j % 4>(g, hist_row, hist_writer, c, p0a, p0b, p0c, (int)p1b.x, p1a.z,
                                                                  lgh + (uint32_t)j, pad);
                    p0a = p1a; p0b = p1b; p1a = fa; p1b = fb;
                    if constexpr (SH) p0c = fc;
                });
            }
        };
        // The short round that opens a dump interval (steps 0..r-1 of the buffer at `cur`, r even, 0 < r < R), software-
        // pipelined like a full round: entered with the records of steps 0 / 1 in (p0, p1) -- the full round in front
        // of it has fetched them -- and left with those of the next round's.  Log bytes go out one by one.
        auto tail_round = [&](auto meas_tag, uint32_t cur, uint32_t nxt, uint32_t lg, int r) {
            constexpr bool MEAS = decltype(meas_tag)::value;
#pragma unroll 1
            for (int j = 0; j < r; ++j) {                 // (rolled: one copy of the step -- this is cold code, and a
                                                          // lone warp pays ~45 cycles for every instruction line it misses)
                uint32_t nrec = rec_addr(cur, j + 2);
                if (j + 2 >= r) nrec = rec_addr(nxt, j + 2 - r);
                if (j + 2 == r) group_barrier<Cfg::GW>(bar_id);
                const uint4 fa = lds_v4(nrec), fb = lds_v4(nrec + Cfg::REC_B);
                uint4 fc = make_uint4(0, 0, 0, 0);
                if constexpr (SH) fc = hist_lookup(p1b.w);
                lat_step<ALGO, LAY, SCSH, SH, SH, MEAS>(g, hist_row, hist_writer, c, p0a, p0b, p0c, (int)p1b.x, p1a.z,
                                                       lg + (uint32_t)j);
                p0a = p1a; p0b = p1b; p1a = fa; p1b = fb;
                if constexpr (SH) p0c = fc;
            }
        };
        // Any other round (segment edges): steps [jb, je), records straight from shared memory.  The
        // kick folded into the last step is the first record of the next round (slot jbn of `nxt`).
        auto edge_round = [&](auto meas_tag, uint32_t cur, uint32_t nxt, uint32_t lg, int jb, int je, int jbn) {
            constexpr bool MEAS = decltype(meas_tag)::value;
            for (int j = jb; j < je; ++j) {
                const uint32_t rec = rec_addr(cur, j);
                const uint4 xa = lds_v4(rec), xb = lds_v4(rec + Cfg::REC_B);
                uint32_t nrec = rec_addr(cur, j + 1);
                if (j == je - 1) { group_barrier<Cfg::GW>(bar_id); nrec = rec_addr(nxt, jbn); }
                const uint4 nb_ = lds_v4(nrec + Cfg::REC_B);
                uint32_t pn = 0u;
                if constexpr (DUAL) pn = lds_u32(nrec + 8u);         // ra.z of the next step's record
                uint4 xc = make_uint4(0, 0, 0, 0);
                if constexpr (SH) xc = hist_lookup(xb.w);
                lat_step<ALGO, LAY, SCSH, SH, SH, MEAS>(g, hist_row, hist_writer, c, xa, xb, xc, (int)nb_.x, pn,
                                                       lg + (uint32_t)j);
            }
        };

        uint32_t cur = 0u, lbuf = 0u;
        const uint32_t dump_full = (uint32_t)(a.dump_every / (uint64_t)R);      // full rounds / steps of the short round
        const int dump_tail = (int)(a.dump_every % (uint64_t)R);                // of a dump interval
        Segments segs(a);
        uint64_t s = a.step_begin;
        bool first = true;
        bool piped = false;                  // (p0, p1) hold the first two records of the coming round
        const bool dump_even = a.dump_every != 0ull && (a.dump_every & 1ull) == 0ull;
        while (s < a.step_end) {
            // A whole dump interval in the steady state: the pause, its full rounds and one short round, nothing else.
            // Everything outside the unrolled full round is cold code for the lone in-order walker (~45 cycles per
            // instruction line it has to fetch, ~25 per branch): the general schedule below cost as much as 20 of the
            // interval's 50 steps.
            if (dump_even && !first && s == segs.next_dump && a.step_end - s >= a.dump_every) {
                segs.next_dump += a.dump_every;
                const uint64_t seg_end = s + a.dump_every;
                PF(pf_seg)
                do_pause();                                          // worm_ising_2d.cpp:120-124
                PF(pf_dump)
                if (!piped) {
                    p0a = lds_v4(rec_addr(cur, 0)); p0b = lds_v4(rec_addr(cur, 0) + Cfg::REC_B);
                    p1a = lds_v4(rec_addr(cur, 1)); p1b = lds_v4(rec_addr(cur, 1) + Cfg::REC_B);
                    if constexpr (SH) {
                        p0c = hist_lookup(p0b.w);
                    }
                    piped = true;
                }
                if (dump_tail) {
                    const uint32_t nxt = cur ^ (uint32_t)Cfg::BUF;
                    tail_round(std::true_type{}, cur, nxt, log_q + lbuf, dump_tail);
                    cur = nxt;
                    lbuf = (lbuf == 2u * Cfg::LOGB) ? 0u : lbuf + (uint32_t)Cfg::LOGB;
                    if constexpr (SH) {
                        if (++flush_ctr == Cfg::FLUSH) { flush_hist(); flush_ctr = 0; }
                    }
                }
                PF(pf_tail)
                for (uint32_t kk = 0; kk < dump_full; ++kk) {
                    const uint32_t nxt = cur ^ (uint32_t)Cfg::BUF;
                    full_round(std::true_type{}, cur, nxt, log_q + lbuf);
                    cur = nxt;
                    lbuf = (lbuf == 2u * Cfg::LOGB) ? 0u : lbuf + (uint32_t)Cfg::LOGB;
                    if constexpr (SH) {
                        if (++flush_ctr == Cfg::FLUSH) { flush_hist(); flush_ctr = 0; }
                    }
                }
                PF(pf_full)
                s = seg_end;
#if WORM_LAT_PROF
                ++pf_n;
#endif
                continue;
            }
            bool meas, dump_first;
            const uint64_t seg_end = segs.next(a, s, meas, dump_first);
            PF(pf_seg)
            if (dump_first) do_pause();
            PF(pf_dump)
            while (s < seg_end) {
                const Batch bt = make_batch<R>(s, seg_end);
                if (first) {
                    group_barrier<Cfg::GW>(bar_id);                  // the first round's records are complete
                    const uint4 kb = lds_v4(rec_addr(cur, bt.jb) + Cfg::REC_B);
                    c.head = c.closed ? (int)kb.x : c.head;          // kick of the first step
                    if constexpr (DUAL) {                            // ... and its bond
                        c.addr = (uint32_t)c.head + lds_u32(rec_addr(cur, bt.jb) + 8u);
                        c.nb = lds_u8(c.addr);
                    }
                    first = false;
                }
                if (bt.jb == 0 && bt.je == R) {
                    if (!piped) {
                        p0a = lds_v4(rec_addr(cur, 0)); p0b = lds_v4(rec_addr(cur, 0) + Cfg::REC_B);
                        p1a = lds_v4(rec_addr(cur, 1)); p1b = lds_v4(rec_addr(cur, 1) + Cfg::REC_B);
                        if constexpr (SH) {
                            p0c = hist_lookup(p0b.w);
                        }
                    }
                    // (one loop per MEAS value, the test at the bottom: the lone walker pays ~30 cycles for every taken
                    // branch, and `if (meas) ... else ...` inside one loop cost it three of them per round)
                    auto rounds = [&](auto meas_tag) {
                        uint32_t left = bt.n;
                        do {
                            const uint32_t nxt = cur ^ (uint32_t)Cfg::BUF;
                            full_round(meas_tag, cur, nxt, log_q + lbuf);
                            cur = nxt;
                            lbuf = (lbuf == 2u * Cfg::LOGB) ? 0u : lbuf + (uint32_t)Cfg::LOGB;
                            if constexpr (SH) {
                                if (++flush_ctr == Cfg::FLUSH) { flush_hist(); flush_ctr = 0; }
                            }
                        } while (--left);
                    };
                    if (meas) rounds(std::true_type{});
                    else rounds(std::false_type{});
                    piped = true;
                    PF(pf_full)
                } else {
                    const uint32_t nxt = cur ^ (uint32_t)Cfg::BUF;
                    if (meas) edge_round(std::true_type{}, cur, nxt, log_q + lbuf, bt.jb, bt.je, (int)(bt.s_next & 1ull));
                    else edge_round(std::false_type{}, cur, nxt, log_q + lbuf, bt.jb, bt.je, (int)(bt.s_next & 1ull));
                    cur = nxt;
                    lbuf = (lbuf == 2u * Cfg::LOGB) ? 0u : lbuf + (uint32_t)Cfg::LOGB;
                    piped = false;
                    if constexpr (SH) {
                        if (++flush_ctr == Cfg::FLUSH) { flush_hist(); flush_ctr = 0; }
                    }
                    PF(pf_edge)
                }
                s = bt.s_next;
            }
#if WORM_LAT_PROF
            ++pf_n;
#endif
        }
#if WORM_LAT_PROF
        if (blockIdx.x == 0 && lane == 0 && w == 0)
            printf("lat prof: segments %lld total %lld cycles: seg %lld dump %lld full %lld tail %lld edge %lld (per segment %lld)\n",
                   pf_n, clock64() - pf_t0, pf_seg, pf_dump, pf_full, pf_tail, pf_edge, (clock64() - pf_t0) / (pf_n ? pf_n : 1));
#endif

        if (first) group_barrier<Cfg::GW>(bar_id);                             // empty step range: meet the producer once
        group_barrier<Cfg::GW>(bar_id);                                        // every step is logged: the accountant may finish
        flush_hist();
        if (live) {
            // a closed worm sits on its tail (head holds the folded kick of the step after the last)
            const int hf = c.closed ? c.tail : c.head;
            a.head[cl] = NIB ? nib_site<CH>(hf) : hf / SC;
            a.tail[cl] = NIB ? nib_site<CH>(c.tail) : c.tail / SC;
            if (a.n_dumps) a.n_dumps[cl] = nd;
        }
    }

    // ---- write back (all but the walker warps, whose spare lanes have retired)
    __syncthreads();
    // participants: every warp but the walkers (CH <= 8: the first (NP+NA) nw warps; solo: warps 1..15)
    const int wb0 = (CH == 32) ? 32 : 0;
    const int pthreads = (CH == 32) ? nthreads - 32 : (NP + NA) * nw * 32;
    if (!is_walker) {
        if constexpr (DUAL) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(smem8);
            const int total = bch * g.N;
            for (int i = tid - wb0; i < total; i += pthreads) {
                const int q = i % CH;
                const int ws = i / CH;
                const int wk = ws / g.N, s = ws - wk * g.N;
                const int64_t cg = chain0 + wk * CH + q;
                if (cg < a.n_chains)
                    reinterpret_cast<uint16_t *>(a.bonds + cg * (int64_t)g.nb2)[s] = (uint16_t)src[i];
            }
        } else if constexpr (NIB) {
            const uint8_t *bad = smem8 + g.bad_off;
            const int total = bch * g.N;
            for (int i = tid - wb0; i < total; i += pthreads) {
                const int q = i % CH;
                const int ws = i / CH;
                const int wk = ws / g.N, s = ws - wk * g.N;
                const int64_t cg = chain0 + wk * CH + q;
                const uint32_t v = smem8[wk * g.region + nib_fmt<CH>(s) + 4 * q];
                if (cg < a.n_chains && !bad[wk * CH + q])
                    reinterpret_cast<uint16_t *>(a.bonds + cg * (int64_t)g.nb2)[s] = (uint16_t)((v & 15u) | ((v >> 4) << 8));
            }
        } else {
            const int vec_per_chain = g.nb2 / 16;
            const int total = bch * vec_per_chain;
            const uint4 *src = reinterpret_cast<const uint4 *>(smem8);
            for (int i = tid - wb0; i < total; i += pthreads) {
                const int cslot = i / vec_per_chain;
                const int64_t cg = chain0 + cslot;
                if (cg < a.n_chains)
                    reinterpret_cast<uint4 *>(a.bonds + cg * (int64_t)g.nb2)[i - cslot * vec_per_chain] = src[i];
            }
        }
    }
}

// ---- launch geometry: layout, warps per block and shared memory for a lattice and CH chains per walker
struct LatPlan {
    bool ok = false;
    int hmode = 0;
    int lay = LAY_DUAL;
    LatGeo g{};
    size_t smem = 0;
};

LatPlan lat_plan_mode(int L, int CH, int algo, bool sh, int smem_optin);

LatPlan lat_plan(int L, int CH, int algo, bool hist, int smem_optin)
{
    // G(r): walker-side bins in shared memory when they fit next to the solo walker's bonds (fewer cycles per step
    // than the accountants' atomics), else the accountants
    const char *hm = getenv("WORM_LAT_HMODE");          // "2": accountants even where the walker-side bins fit (experiments)
    if (hist && CH == 32 && !(hm && hm[0] == '2')) {
        LatPlan p = lat_plan_mode(L, CH, algo, true, smem_optin);
        if (p.ok) { p.hmode = 1; return p; }
    }
    LatPlan p = lat_plan_mode(L, CH, algo, false, smem_optin);
    p.hmode = hist ? 2 : 0;
    return p;
}

LatPlan lat_plan_mode(int L, int CH, int algo, bool sh, int smem_optin)
{
    LatPlan p;
    if (L < 4 || (L & (L - 1)) || (CH != 32 && CH != 8 && CH != 4 && CH != 1)) return p;
    const int N = L * L;
    const int NP = (CH == 32) ? 8 : (CH == 8) ? 2 : 1;
    const int R = 32 * NP * 2 / CH;
    const size_t histb = sh ? (size_t)CH * N * 2 + 64 : 0;        // 16-bit bins + the four displacement-update entries
    const size_t recw = (size_t)4 * R * CH * 16;                  // LatCfg::RECW
    const size_t logw = (size_t)3 * (R + 4) * CH;                 // LatCfg::LOGW: three rounds of one byte per step
    // Throughput = (chains in flight) / (cycles per step): of the (layout, walkers per block) pairs that
    // fit, take the one with the most chains per block; ties go to the layout with the fewest instructions in
    // front of the bond load (dual, then byte per bond, then the nibble layout, which masks its occupation).
    // WORM_LAT_LAYOUT=dual|byte|nibble pins the layout (experiments).
    const char *pin = getenv("WORM_LAT_LAYOUT");
    int best = 0;
    for (int lay : {LAY_DUAL, LAY_BYTE, LAY_NIB}) {
        if (lay == LAY_BYTE && CH != 1) continue;
        if (lay == LAY_NIB && algo != ALGO_TAYLOR) continue;
        if (pin && ((pin[0] == 'd') != (lay == LAY_DUAL) || (pin[0] == 'b') != (lay == LAY_BYTE) || (pin[0] == 'n') != (lay == LAY_NIB))) continue;
        const size_t region = lay == LAY_DUAL ? (size_t)4 * N * CH : lay == LAY_BYTE ? (size_t)2 * N : (size_t)N * CH;
        const int sc = lay == LAY_DUAL ? 4 * CH : lay == LAY_BYTE ? 2 : CH;
        const char *nwcap = getenv("WORM_LAT_NW");         // cap on the walker groups per block (experiments)
        for (int nw : {4, 3, 2, 1}) {
            if (CH == 32 && nw != 1) continue;
            if (nwcap && nw > nwcap[0] - '0') continue;
            // Groups of four chains: several one-group blocks per SM beat one block of several groups when as many
            // chains fit that way (L = 128: three 69 KB blocks, 75 cycles per step against 89 for one block of three
            // groups).  Not so for eight chains per group (L = 64: 102 against 224) or one (L = 256: 94 against 163).
            if (CH == 4 && nw > 1 && lay != LAY_DUAL) {
                const size_t need1 = ((region + recw + logw) + 15) / 16 * 16 + 16;
                if ((size_t)(228 * 1024) / (need1 + 1024) >= (size_t)nw) continue;
            }
            if (nw * CH <= best) continue;
            const size_t body = (nw * (region + recw + logw) + 15) / 16 * 16 + histb;
            const size_t need = body + (lay == LAY_NIB ? (size_t)((nw * CH + 15) / 16 * 16) : 0);
            if (need + 1024 > (size_t)smem_optin) continue;
            best = nw * CH;
            p.ok = true; p.lay = lay; p.smem = need;
            p.g.L = L; p.g.N = N; p.g.nb2 = 2 * N; p.g.nw = nw;
            p.g.SC = sc;
            p.g.XM = (L - 1) * p.g.SC;
            p.g.YM = (L - 1) * L * p.g.SC;
            p.g.XP = p.g.SC; p.g.XN = -p.g.SC;
            if (lay == LAY_NIB) {            // column format (nib_fmt): x = bits 0-1 and the bits above the column index
                p.g.XM = ((L / 4 - 1) * 4 * CH) | 3;
                p.g.XP = 1 + 4 * (CH - 1);
                p.g.XN = -1;
            }
            p.g.region = (int)region;
            p.g.rec_off = (int)(nw * region);
            p.g.log_off = (int)(nw * (region + recw));
            p.g.hist_off = (int)((nw * (region + recw + logw) + 15) / 16 * 16);
            p.g.htab_off = p.g.hist_off + (int)((size_t)CH * N * 2);
            p.g.bad_off = (int)body;
            p.g.HXF = 2 | ((L / 2 - 1) << 7);
            p.g.HYF = (L - 1) * L * 64;
            p.g.lowt = 0;
            p.g.zero = 0;
        }
    }
    return p;
}

template <int ALGO, int CH, int LAY, int HMODE>
cudaError_t launch_lat(const PhiloxArgs &a, const LatPlan &p, cudaStream_t st)
{
    auto kern = worm_lat_kernel<ALGO, CH, LAY, HMODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    const int64_t bch = (int64_t)p.g.nw * CH;
    const int64_t blocks = (a.n_chains + bch - 1) / bch;
    const int warps = (CH == 32) ? 16 : LatCfg<CH, false>::GW * p.g.nw;
    kern<<<(unsigned)blocks, warps * 32, p.smem, st>>>(a, p.g);
    return cudaGetLastError();
}

template <int ALGO, int CH, int LAY>
cudaError_t launch_lat_h(const PhiloxArgs &a, const LatPlan &p, cudaStream_t st)
{
    if (p.hmode == 0) return launch_lat<ALGO, CH, LAY, 0>(a, p, st);
    if (p.hmode == 2) return launch_lat<ALGO, CH, LAY, 2>(a, p, st);
    if constexpr (CH == 32) return launch_lat<ALGO, CH, LAY, 1>(a, p, st);
    return cudaErrorInvalidValue;
}

template <int ALGO>
cudaError_t launch_lat_a(const PhiloxArgs &a, const LatPlan &p, int CH, cudaStream_t st)
{
    if (p.lay == LAY_BYTE) return launch_lat_h<ALGO, 1, LAY_BYTE>(a, p, st);
    if (p.lay == LAY_NIB) {
        if constexpr (ALGO == ALGO_TAYLOR) {
            switch (CH) {
            case 32: return launch_lat_h<ALGO, 32, LAY_NIB>(a, p, st);
            case 8: return launch_lat_h<ALGO, 8, LAY_NIB>(a, p, st);
            case 4: return launch_lat_h<ALGO, 4, LAY_NIB>(a, p, st);
            case 1: return launch_lat_h<ALGO, 1, LAY_NIB>(a, p, st);
            default: return cudaErrorInvalidValue;
            }
        } else {
            return cudaErrorInvalidValue;
        }
    }
    switch (CH) {
    case 32: return launch_lat_h<ALGO, 32, LAY_DUAL>(a, p, st);
    case 8: return launch_lat_h<ALGO, 8, LAY_DUAL>(a, p, st);
    case 4: return launch_lat_h<ALGO, 4, LAY_DUAL>(a, p, st);
    case 1: return launch_lat_h<ALGO, 1, LAY_DUAL>(a, p, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// G = lanes_per_chain of the C ABI: 2 / 4 / 8 / 32 <-> 32 / 8 / 4 / 1 chains per walker warp
static int chains_per_walker(int G) { return G == 2 ? 32 : 32 / G; }

int lat_chains_per_block(int L, int G, int algo, bool hist, int smem_optin)
{
    const LatPlan p = lat_plan(L, chains_per_walker(G), algo, hist, smem_optin);
    return p.ok ? p.g.nw * chains_per_walker(G) : 0;
}

bool lat_info(int L, int G, int algo, bool hist, int smem_optin, int *chains_per_block, int *walkers, int *layout)
{
    if (G != 2 && G != 4 && G != 8 && G != 32) return false;
    const LatPlan p = lat_plan(L, chains_per_walker(G), algo, hist, smem_optin);
    if (!p.ok) return false;
    // blocks that fit an SM together (shared memory incl. the 1 KB the system keeps per block; threads)
    const int warps = (chains_per_walker(G) == 32) ? 16 : (chains_per_walker(G) == 8 ? 4 : 3) * p.g.nw;
    int per_sm = (int)((size_t)(228 * 1024) / (p.smem + 1024));
    if (per_sm > 64 / warps) per_sm = 64 / warps;
    if (per_sm < 1) per_sm = 1;
    if (chains_per_walker(G) != 4 || p.g.nw != 1) per_sm = 1;      // only the one-group blocks of four chains are run that way
    *chains_per_block = p.g.nw * chains_per_walker(G) * per_sm;    // chains an SM holds
    *walkers = p.g.nw * per_sm;
    *layout = p.lay;
    return true;
}

size_t lat_smem_bytes(int L, int G, int algo, bool hist, int smem_optin)
{
    if (G != 2 && G != 4 && G != 8 && G != 32) return 0;
    const LatPlan p = lat_plan(L, chains_per_walker(G), algo, hist, smem_optin);
    return p.ok ? p.smem : 0;
}

cudaError_t launch_philox_lat(const PhiloxArgs &a, int G, bool lowt, int smem_optin, cudaStream_t st)
{
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    const int CH = chains_per_walker(G);
    LatPlan p = lat_plan(a.L, CH, a.algo, hist, smem_optin);
    if (!p.ok) return cudaErrorInvalidValue;
    p.g.lowt = lowt ? 1 : 0;
    if (a.algo == ALGO_TAYLOR) return launch_lat_a<ALGO_TAYLOR>(a, p, CH, st);
    return launch_lat_a<ALGO_BINARY>(a, p, CH, st);
}

}  // namespace worm
