// worm_pack.cu -- production worm kernel, "packed throughput mode" (sm_100a): one THREAD per Markov
// chain, bond occupations packed in shared memory.
//
// For batches with (many) more chains than the GPU has issue slots, and for lattices whose byte-per-bond
// state leaves the latency kernel (worm_lat.cu) with a handful of chains per SM.  Throughput of a batch is
//     min( chains in flight / cycles per step ,  issue slots / instructions per step ),
// so the kernel is built around (1) bytes of shared memory per chain and (2) thread-instructions per step:
//
//  * Packed state (the rule being replaced is worm_ising_2d.cpp:139-155, which reads and writes one
//    `int bonds[]` element per step).  Taylor chain: 4 bits per bond (occupations 0..15; an occupation
//    that would reach 16 is reported in the chain's WORM_ACC_STATUS word, never wrapped silently), i.e.
//    1 KiB per chain at L = 32 -> 224 chains per SM instead of 32.  Binary chain: 1 bit per bond, 256 B
//    per chain at L = 32 (the north star's layout) -> up to 512 chains per SM (128 registers per thread).
//    Word w of the chain of lane l lives at region(warp) + (w * 32 + l) * 4: every lane of a warp owns
//    its own bank, so the random per-lane accesses of a step never conflict (a ragged last warp packs its
//    columns densely).
//  * A step is ~63 (Taylor) / ~51 (binary) thread-instructions, branch-free.  The head is a site index scaled to the BIT offset of
//    the site's +x bond inside the chain's packed array (x 8 Taylor, x 2 binary): a move is IADD + LOP3
//    (x / y bit fields wrap by masking, power-of-two L), the bond's word is `offset >> 5`, its shift the low
//    five bits.  Everything a step needs that depends only on the random word's top four bits
//    (direction, +/- coin) -- head increment, wrap field, "bond owned by the neighbour" mask merged with
//    the +y offset, occupation change -- comes out of an 8-entry shared-memory table (eight conflict-free
//    copies of each entry) with one SHF + one LOP3 + one LDS.128, instead of a chain of predicates and selects.
//  * The Philox block and the table entries of the NEXT pair of steps are drawn one pair ahead (two pairs per
//    loop trip, register sets alternating): they do not depend on the chain's state, so their instructions fill
//    the issue slots the dependent step chain (address -> LDS -> accept -> head) leaves empty.  Round keys live
//    in registers.
//  * Observables in registers: Z, sum Nb, sum n as 32-bit partial sums folded into 64-bit totals every
//    `fold` steps, sum Nb^2 / sum n^2 as one IMAD.WIDE each, all predicated on the closed test.
//  * G(r): the displacement head - tail is carried along in the same scaled format (it IS the byte offset of
//    its int64 bin); iterations at an unchanged displacement are counted in a register and added with ONE
//    `red.global.add.u64` when an accepted move changes it.  Resident blocks come from all over the batch
//    (PackGeo::perm), so the atomics of a moment spread over all temperatures' rows.
//  * Dumps (worm_ising_2d.cpp:113-125): the step range is cut at dump multiples by uniform loop control only;
//    the lanes of a warp copy the configuration of every chain that dumps, unpacking as they go.
//
// Algorithm and random stream: oracle/worm_oracle.c:worm_oracle_philox_run (bit-for-bit), i.e. the
// reference loop worm_ising_2d.cpp:111-157 on a Philox4x32-10 stream.  Power-of-two L >= 4.
#include "worm_run.cuh"

#ifndef WORM_PACK_BIG
#define WORM_PACK_BIG 512        // threads per block of the high-occupancy instantiation (registers: 65536 / this)
#endif
#ifndef WORM_PACK_UNROLL2
#define WORM_PACK_UNROLL2 1      // two pairs of steps per loop trip (register sets alternate) instead of one pair + copies
#endif

namespace worm {

namespace {

struct PackGeo {
    int L, N, nb2;
    int W;               // 32-bit words of packed bonds per chain
    int CB;              // chains per block
    int TPB;             // threads per block (CB rounded up to a whole warp)
    int ksh;             // kick: scaled site = (b >> ksh) & kmask
    uint32_t kmask;
    uint32_t fold;       // steps between folds of the 32-bit partial sums (even)
    uint32_t perm;       // block b walks chain group (b * perm) % gridDim.x: blocks resident together come from
                         // all over the batch (all temperatures of a temperature-major batch at once, so the G(r)
                         // atomics of a moment spread over all histogram rows instead of hammering one)
};

// Step table: 8 entries (direction, coin) of {add, field, own | y offset, occupation change}, each in eight
// copies 16 bytes apart -- lane l reads copy l % 8, so the eight lanes of a 128-bit load phase always hit
// eight different bank quads whatever entries they want (no conflicts).  Entry i at byte i * 128.
constexpr int LUT_BYTES = 1024;
__device__ __forceinline__ uint32_t lut_offset(uint32_t rb, uint32_t lane_off)
{
    return ((rb >> 22) & 0x380u) | lane_off;
}

template <int ALGO> struct PackFmt {
    static constexpr int US = (ALGO == ALGO_TAYLOR) ? 3 : 1;          // log2(bits per site)
    static constexpr uint32_t YOFF = (ALGO == ALGO_TAYLOR) ? 4u : 1u; // bit offset of the +y bond within a site
    static constexpr uint32_t SITEBITS = 31u & ~((1u << US) - 1u);    // bits of a shift that come from the site
    static constexpr uint32_t OCCMASK = (ALGO == ALGO_TAYLOR) ? 15u : 1u;
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct PackChain {
    uint32_t hs, tail;           // scaled site indices (site << US)
    uint32_t disp;               // HIST: head - tail, same scaling
    uint32_t run;                // HIST: measured iterations spent at this displacement and not yet counted
    int Nb, npar;                // sum of occupations, number of odd bonds
    uint32_t Zp, SNbp, Snp;      // partial sums since the last fold
    uint64_t Z, SNb, SNb2, Sn, Sn2;
};

struct PackConst {
    uint32_t base;               // shared address of word 0 of this thread's chain
    uint32_t pitch;              // bytes between consecutive words of a chain (128 in a whole warp, 4 x its columns in a ragged one)
    uint32_t lut;                // shared address of the step table
    uint32_t lut_lane;           // byte offset of this lane's copy of a table entry
    uint32_t klo, khi;           // kfix (Taylor) / hfix (binary), 64 bits
    uint32_t tlo, thi;           // tfix (Taylor, T < 1 only)
    unsigned long long hist_row; // HIST: this chain's row of the int64 histogram
    uint32_t hist_inc;           // HIST: 1 for a live chain, 0 for the lanes that only shadow one
    int ksh;
    uint32_t kmask;
};

template <int ALGO>
__device__ __forceinline__ void pack_hist_flush(const PackConst &k, PackChain &c)
{
    const unsigned long long p = k.hist_row + (unsigned long long)c.disp * (ALGO == ALGO_TAYLOR ? 1u : 4u);
    asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"((unsigned long long)c.run) : "memory");
    c.run = 0u;
}

// One worm step, worm_ising_2d.cpp:112-156 on the words (a, b) of oracle/worm_oracle.c's stream.
// e: the step-table entry of rb's (direction, coin), fetched ahead of time.
template <int ALGO, bool LOWT, bool MEAS, bool HIST>
__device__ __forceinline__ void pack_step(const PackConst &k, PackChain &c, const uint32_t ra, const uint32_t rb, const uint4 e)
{
    using F = PackFmt<ALGO>;
    const bool closed = (c.hs == c.tail);
    if constexpr (MEAS) {
        if (closed) {                                              // :130-132
            c.Zp += 1u;
            c.SNbp += (uint32_t)c.Nb;
            c.SNb2 += (uint64_t)(uint32_t)c.Nb * (uint32_t)c.Nb;
            if constexpr (ALGO == ALGO_TAYLOR) {
                c.Snp += (uint32_t)c.npar;
                c.Sn2 += (uint64_t)(uint32_t)c.npar * (uint32_t)c.npar;
            }
        }
        // G(r): this iteration is one more at the current displacement head - tail (taken before the kick);
        // the count goes to the bin when the displacement changes (below) or the range ends
        if constexpr (HIST) c.run += k.hist_inc;
    }
    const uint32_t kick = (rb >> k.ksh) & k.kmask;                 // :128-129
    const uint32_t hs = closed ? kick : c.hs;
    c.tail = closed ? kick : c.tail;
    // direction :137, neighbour :88-95, bond_number :214-244 through the step table
    const uint32_t t = hs + e.x;
    const uint32_t nh = (hs & ~e.y) | (t & e.y);
    const uint32_t own = (nh & e.z) | (hs & ~e.z);                 // bit offset of the site that owns the bond
    const uint32_t sh = (own & F::SITEBITS) | (e.z & ~F::SITEBITS);// + y offset (high bits: ignored by the wrap shifts)
    const uint32_t addr = k.base + (own >> 5) * k.pitch;
    const uint32_t word = lds32(addr);                             // :139
    const uint32_t v = __funnelshift_r(word, 0u, sh);
    uint32_t dn = 0u;
    if constexpr (HIST) {
        const uint32_t td = c.disp + e.x;
        dn = (c.disp & ~e.y) | (td & e.y);
    }
    // Accept test (:141-151) and update (:152-154) in one PTX block: the three-way predicate
    // acc = dec ? a_dec : a_inc stays a predicate (one PLOP3) instead of a chain of selects, and the
    // store, the counters and the head all hang off it.
    uint32_t hs_new, disp_new = 0u;
    if constexpr (ALGO == ALGO_TAYLOR) {
        const uint32_t neww = word + __funnelshift_l(0u, e.w, sh); // occupation +/- 1 (a carry out of 15 is caught at the end)
        if constexpr (!LOWT) {
            asm volatile("{\n\t"
                         ".reg .pred pd, pn, pi, pa;\n\t"
                         ".reg .b32 nb, m1, t2;\n\t"
                         ".reg .b64 pr, kk;\n\t"
                         "and.b32 nb, %5, 15;\n\t"
                         "setp.lt.s32 pd, %6, 0;\n\t"                // decrement proposal :141
                         "setp.ne.u32 pn, nb, 0;\n\t"                // u < nb/K  (T >= 1: nb/K is 0 or >= 1)  :148,151
                         "add.u32 m1, nb, 1;\n\t"
                         "mul.wide.u32 pr, m1, %7;\n\t"
                         "mov.b64 kk, {%8, %9};\n\t"
                         "setp.lt.u64 pi, pr, kk;\n\t"               // u < K/(nb+1)  :143,151
                         "and.pred pn, pn, pd;\n\t"
                         "not.pred pd, pd;\n\t"
                         "and.pred pi, pi, pd;\n\t"
                         "or.pred pa, pn, pi;\n\t"
                         "@pa st.shared.u32 [%10], %11;\n\t"         // :152
                         "@pa add.s32 %0, %0, %6;\n\t"               // :153
                         "shl.b32 t2, %5, 1;\n\t"
                         "and.b32 t2, t2, 2;\n\t"
                         "sub.s32 t2, 1, t2;\n\t"
                         "@pa add.s32 %1, %1, t2;\n\t"
                         "selp.b32 %2, %12, %13, pa;\n\t"            // :154
                         "selp.b32 %3, %14, %4, pa;\n\t"
                         "}"
                         : "+r"(c.Nb), "+r"(c.npar), "=r"(hs_new), "=r"(disp_new)
                         : "r"(c.disp), "r"(v), "r"(e.w), "r"(ra), "r"(k.klo), "r"(k.khi), "r"(addr), "r"(neww), "r"(nh), "r"(hs),
                           "r"(dn)
                         : "memory");
        } else {
            asm volatile("{\n\t"
                         ".reg .pred pd, pn, pi, pa;\n\t"
                         ".reg .b32 nb, m1, t2, ql, qh, x;\n\t"
                         ".reg .b64 pr, kk, q, r64;\n\t"
                         "and.b32 nb, %5, 15;\n\t"
                         "setp.lt.s32 pd, %6, 0;\n\t"
                         "mul.wide.u32 q, nb, %15;\n\t"              // nb * tfix (64 bits)
                         "mov.b64 {ql, qh}, q;\n\t"
                         "mad.lo.u32 qh, nb, %16, qh;\n\t"
                         "mov.b64 q, {ql, qh};\n\t"
                         "cvt.u64.u32 r64, %7;\n\t"
                         "setp.lt.u64 pn, r64, q;\n\t"               // u < nb/K  :148,151
                         "add.u32 m1, nb, 1;\n\t"
                         "mul.wide.u32 pr, m1, %7;\n\t"
                         "mov.b64 kk, {%8, %9};\n\t"
                         "setp.lt.u64 pi, pr, kk;\n\t"
                         "and.pred pn, pn, pd;\n\t"
                         "not.pred pd, pd;\n\t"
                         "and.pred pi, pi, pd;\n\t"
                         "or.pred pa, pn, pi;\n\t"
                         "@pa st.shared.u32 [%10], %11;\n\t"
                         "@pa add.s32 %0, %0, %6;\n\t"
                         "shl.b32 t2, %5, 1;\n\t"
                         "and.b32 t2, t2, 2;\n\t"
                         "sub.s32 t2, 1, t2;\n\t"
                         "@pa add.s32 %1, %1, t2;\n\t"
                         "selp.b32 %2, %12, %13, pa;\n\t"
                         "selp.b32 %3, %14, %4, pa;\n\t"
                         "}"
                         : "+r"(c.Nb), "+r"(c.npar), "=r"(hs_new), "=r"(disp_new)
                         : "r"(c.disp), "r"(v), "r"(e.w), "r"(ra), "r"(k.klo), "r"(k.khi), "r"(addr), "r"(neww), "r"(nh), "r"(hs),
                           "r"(dn), "r"(k.tlo), "r"(k.thi)
                         : "memory");
        }
    } else {
        const uint32_t neww = word ^ __funnelshift_l(0u, 1u, sh);
        // an occupied bond is always taken, an empty one with probability tanh K (hfix; khi != 0: T so low that it is 1)
        asm volatile("{\n\t"
                     ".reg .pred pn, pi, ph, pa;\n\t"
                     ".reg .b32 nb, t2;\n\t"
                     "and.b32 nb, %4, 1;\n\t"
                     "setp.ne.u32 pn, nb, 0;\n\t"
                     "setp.lt.u32 pi, %5, %6;\n\t"
                     "setp.ne.u32 ph, %7, 0;\n\t"
                     "or.pred pa, pn, pi;\n\t"
                     "or.pred pa, pa, ph;\n\t"
                     "@pa st.shared.u32 [%8], %9;\n\t"
                     "shl.b32 t2, nb, 1;\n\t"
                     "sub.s32 t2, 1, t2;\n\t"
                     "@pa add.s32 %0, %0, t2;\n\t"
                     "selp.b32 %1, %10, %11, pa;\n\t"
                     "selp.b32 %2, %12, %3, pa;\n\t"
                     "}"
                     : "+r"(c.Nb), "=r"(hs_new), "=r"(disp_new)
                     : "r"(c.disp), "r"(v), "r"(ra), "r"(k.klo), "r"(LOWT ? k.khi : 0u), "r"(addr), "r"(neww), "r"(nh), "r"(hs), "r"(dn)
                     : "memory");
    }
    c.hs = hs_new;
    if constexpr (HIST) {
        // An accepted move always changes the head (L >= 3), hence the displacement: the run of iterations
        // counted at the old one goes to its bin.  The scaled displacement is the byte offset of its int64 bin
        // (Taylor: site * 8) or a quarter of it (binary: site * 2).  One predicated `red` per accepted move
        // instead of one per step, and far fewer hits on the hot bins (a closed worm waits in bin 0).
        // Lanes without a chain count nothing (hist_inc = 0): their `red` adds 0 to a valid bin.
        if constexpr (MEAS) {
            const unsigned long long p = k.hist_row + (unsigned long long)c.disp * (ALGO == ALGO_TAYLOR ? 1u : 4u);
            uint32_t run_new;
            asm volatile("{\n\t"
                         ".reg .pred p;\n\t"
                         ".reg .b64 rr;\n\t"
                         "setp.ne.u32 p, %1, %2;\n\t"
                         "mov.b64 rr, {%3, %4};\n\t"
                         "@p red.global.add.u64 [%5], rr;\n\t"
                         "selp.b32 %0, 0, %3, p;\n\t"
                         "}"
                         : "=r"(run_new) : "r"(hs_new), "r"(hs), "r"(c.run), "r"(0u), "l"(p) : "memory");
            c.run = run_new;
        }
        c.disp = disp_new;
    }
}

__device__ __forceinline__ void pack_fold(PackChain &c)
{
    c.Z += c.Zp; c.SNb += c.SNbp; c.Sn += c.Snp;
    c.Zp = c.SNbp = c.Snp = 0u;
}

// ---- packing: uint8 [2N] (bond 2*site = +x, 2*site+1 = +y) <-> 4 / 1 bits per bond, in bond order
__device__ __forceinline__ uint32_t bit16(const uint4 b, uint32_t &bad)   // sixteen 0/1 occupations (bytes) -> 16 bits
{
    bad |= (b.x | b.y | b.z | b.w) & 0xfefefefeu;
    const uint32_t n0 = ((b.x & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t n1 = ((b.y & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t n2 = ((b.z & 0x01010101u) * 0x01020408u) >> 24;
    const uint32_t n3 = ((b.w & 0x01010101u) * 0x01020408u) >> 24;
    return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}
__device__ __forceinline__ uint32_t nib4(uint32_t x)         // four occupations (bytes) -> 16 bits
{
    const uint32_t y = (x | (x >> 4)) & 0x00ff00ffu;
    return (y | (y >> 8)) & 0xffffu;
}
__device__ __forceinline__ uint32_t unnib4(uint32_t h)       // 16 bits -> four occupations (bytes)
{
    const uint32_t y = (h | (h << 8)) & 0x00ff00ffu;
    return (y | (y << 4)) & 0x0f0f0f0fu;
}
__device__ __forceinline__ uint32_t unbit4(uint32_t n)       // 4 bits -> four occupations (bytes)
{
    return ((n & 15u) * 0x00204081u) & 0x01010101u;
}

// the packed words [w0, w0 + nwords) of one chain, read from its column in shared memory and written to a
// uint8 [2N] array in global memory by the calling lane
template <int ALGO>
__device__ __forceinline__ void unpack_word(uint32_t word, uint8_t *dst, int w)
{
    if constexpr (ALGO == ALGO_TAYLOR) {
        uint2 o;
        o.x = unnib4(word & 0xffffu);
        o.y = unnib4(word >> 16);
        reinterpret_cast<uint2 *>(dst)[w] = o;
    } else {
        uint4 o0, o1;
        o0.x = unbit4(word); o0.y = unbit4(word >> 4); o0.z = unbit4(word >> 8); o0.w = unbit4(word >> 12);
        o1.x = unbit4(word >> 16); o1.y = unbit4(word >> 20); o1.z = unbit4(word >> 24); o1.w = unbit4(word >> 28);
        reinterpret_cast<uint4 *>(dst)[2 * w] = o0;
        reinterpret_cast<uint4 *>(dst)[2 * w + 1] = o1;
    }
}

// MAXT: 256 (all the registers a thread can use) or 512 (128 registers) threads per block
template <int ALGO, bool LOWT, bool HIST, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
worm_pack_kernel(const PhiloxArgs a, const PackGeo g)
{
    using F = PackFmt<ALGO>;
    extern __shared__ __align__(16) uint32_t psm[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(psm);

    // ---- step table: entry i = (direction d = i >> 1, coin = i & 1), eight copies each
    for (int slot = tid; slot < 64; slot += (int)blockDim.x) {
        const uint32_t i = (uint32_t)slot >> 3;
        const uint32_t d = i >> 1, coin = i & 1u;
        const bool ym = d & 1u, neg = d & 2u;
        const uint32_t unit = 1u << F::US;
        const uint32_t stp = ym ? unit * (uint32_t)g.L : unit;
        const uint32_t xf = (uint32_t)(g.L - 1) << F::US, yf = xf * (uint32_t)g.L;
        uint4 e;
        e.x = neg ? (0u - stp) : stp;                              // head increment
        e.y = ym ? yf : xf;                                        // the field that moves (and wraps)
        e.z = (neg ? ~(unit - 1u) : 0u) | (ym ? F::YOFF : 0u);     // bond owned by the neighbour | +y offset
        e.w = coin ? 0xffffffffu : 1u;                             // occupation change (Taylor)
        reinterpret_cast<uint4 *>(psm)[slot] = e;
    }
    __syncthreads();
    // Columns: thread t walks the chain in column t.  The columns of a whole warp are 32 words apart
    // (every lane its own bank); a last, partial warp packs its columns densely (its lanes may meet in a
    // bank now and then -- a cycle of the shared-memory pipe, which is idle most of the time).
    const int wfirst = tid & ~31;                                  // first column of this warp
    const int wcols = min(32, g.CB - wfirst);                      // columns of this warp (>= 1)
    const bool has_col = lane < wcols;
    const unsigned wmask = __ballot_sync(0xffffffffu, has_col);
    const int wlanes = wcols;

    const int64_t grp = (int64_t)(((unsigned long long)blockIdx.x * g.perm) % gridDim.x);
    const int64_t cidx = grp * g.CB + tid;
    const bool live = has_col && cidx < a.n_chains;
    const int64_t cl = live ? cidx : 0;          // dead lanes shadow chain 0 of the launch in their own column, never write
    const PhiloxChain prm = a.prm[cl];

    PackConst k;
    k.pitch = (uint32_t)wcols * 4u;
    k.base = sbase + LUT_BYTES + (uint32_t)wfirst * (uint32_t)g.W * 4u + (uint32_t)lane * 4u;
    k.lut = sbase;
    k.lut_lane = (uint32_t)(lane & 7) * 16u;
    if constexpr (ALGO == ALGO_TAYLOR) { k.klo = (uint32_t)prm.kfix; k.khi = (uint32_t)(prm.kfix >> 32); }
    else { k.klo = (uint32_t)prm.hfix; k.khi = (uint32_t)(prm.hfix >> 32); }
    k.tlo = (uint32_t)prm.tfix; k.thi = (uint32_t)(prm.tfix >> 32);
    k.hist_row = 0ull;
    k.hist_inc = 0u;
    if constexpr (HIST) {
        k.hist_row = (unsigned long long)(a.hist + (int64_t)a.group_index[cl] * g.N);
        k.hist_inc = live ? 1u : 0u;
    }
    k.ksh = g.ksh; k.kmask = g.kmask;

    // ---- stage the bonds, packed.  All 32 lanes of a warp read one chain's uint8 array together
    // (coalesced 16-byte loads), pack, and store into that chain's column (one bank: the stores serialise
    // in the shared-memory pipe, which has nothing else to do yet); four chains in flight per pass.  Sum of
    // occupations, number of odd bonds and the "does not fit the packed format" flag are reduced per chain.
    uint32_t bad = 0u;
    PackChain c;
    {
        constexpr int WPP = (ALGO == ALGO_TAYLOR) ? 16 : 32;   // bytes of the uint8 array per "piece": two words / one word
        const int pieces = g.nb2 / WPP;
        int myNb = 0, mynpar = 0;
        for (int j0 = 0; j0 < wcols; j0 += 4) {
            int sNb[4] = {0, 0, 0, 0}, snp[4] = {0, 0, 0, 0};
            uint32_t sbad[4] = {0u, 0u, 0u, 0u};
            long long chain[4];
            bool cok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;
                chain[u] = __shfl_sync(0xffffffffu, (long long)cl, j < wcols ? j : 0);
                cok[u] = j < wcols;
            }
            for (int p0 = 0; p0 < pieces; p0 += 32) {
                const int pc = p0 + lane;
                uint4 v[4], v2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = v2[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (cok[u] && pc < pieces) {
                        const uint4 *src = reinterpret_cast<const uint4 *>(a.bonds + chain[u] * (long long)g.nb2);
                        if constexpr (ALGO == ALGO_TAYLOR) v[u] = src[pc];
                        else { v[u] = src[2 * pc]; v2[u] = src[2 * pc + 1]; }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!cok[u] || pc >= pieces) continue;
                    const uint32_t col = k.base + (uint32_t)(j0 + u - lane) * 4u;
                    const uint4 bb = v[u];
                    if constexpr (ALGO == ALGO_TAYLOR) {
                        sbad[u] |= (bb.x | bb.y | bb.z | bb.w) & 0xf0f0f0f0u;
                        const uint32_t w0 = nib4(bb.x) | (nib4(bb.y) << 16);
                        const uint32_t w1 = nib4(bb.z) | (nib4(bb.w) << 16);
                        sts32(col + (uint32_t)(2 * pc) * k.pitch, w0);
                        sts32(col + (uint32_t)(2 * pc + 1) * k.pitch, w1);
                        sNb[u] += (int)(__vsadu4(bb.x, 0u) + __vsadu4(bb.y, 0u) + __vsadu4(bb.z, 0u) + __vsadu4(bb.w, 0u));
                        snp[u] += __popc(w0 & 0x11111111u) + __popc(w1 & 0x11111111u);
                    } else {
                        const uint32_t w0 = bit16(bb, sbad[u]) | (bit16(v2[u], sbad[u]) << 16);
                        sts32(col + (uint32_t)pc * k.pitch, w0);
                        sNb[u] += __popc(w0);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sNb[u] += __shfl_xor_sync(0xffffffffu, sNb[u], o);
                    snp[u] += __shfl_xor_sync(0xffffffffu, snp[u], o);
                    sbad[u] |= __shfl_xor_sync(0xffffffffu, sbad[u], o);
                }
                if (lane == j0 + u) { myNb = sNb[u]; mynpar = snp[u]; bad = sbad[u]; }
            }
        }
        c.Nb = myNb; c.npar = (ALGO == ALGO_TAYLOR) ? mynpar : myNb;
    }
    int64_t *ac = a.acc + cl * ACC_STRIDE;
    if (bad) {
        // an occupation the packed format cannot hold: report, leave the chain untouched
        if (live) ac[ACC_STATUS] |= 1;
    }
    // a block whose chains cannot all be packed still has to walk the others: the bad chain walks its
    // (truncated) copy and is not written back
    const bool writable = live && !bad;

    c.hs = (uint32_t)a.head[cl] << F::US;
    c.tail = (uint32_t)a.tail[cl] << F::US;
    {
        const uint32_t xf = (uint32_t)(g.L - 1) << F::US, yf = xf * (uint32_t)g.L;
        const uint32_t t1 = c.hs - c.tail;
        const uint32_t t2 = (c.hs | xf) - (c.tail & yf);           // y fields, no borrow from x
        c.disp = (t1 & xf) | (t2 & yf);
    }
    c.run = 0u;
    c.Zp = c.SNbp = c.Snp = 0u;
    c.Z = (uint64_t)ac[0]; c.SNb = (uint64_t)ac[1]; c.SNb2 = (uint64_t)ac[2];
    c.Sn = (uint64_t)ac[3]; c.Sn2 = (uint64_t)ac[4];
    const uint64_t old0 = c.Z, old1 = c.SNb, old2 = c.SNb2, old3 = c.Sn, old4 = c.Sn2;
    int32_t nd = (a.n_dumps && live) ? a.n_dumps[cl] : 0;

    const uint64_t chain_id = a.chain_id0 + (uint64_t)cl;
    const uint32_t c2 = (uint32_t)chain_id, c3 = (uint32_t)(chain_id >> 32);
    const PhiloxKeys keys = philox_keys((uint32_t)prm.seed, (uint32_t)(prm.seed >> 32));
    __syncwarp();

    // steps [s, e) with MEAS constant, no dump point inside and all in one 2^33-step window (the high
    // word of the Philox pair counter is constant: the loop counts in 32 bits)
    auto table = [&](uint32_t rb) { return lds128(k.lut + lut_offset(rb, k.lut_lane)); };
    auto run_range = [&](auto meas_tag, const uint64_t s0, const uint64_t e0) {
        constexpr bool MEAS = decltype(meas_tag)::value;
        const uint32_t phi = (uint32_t)(s0 >> 33);
        uint32_t plo = (uint32_t)(s0 >> 1);
        uint32_t left = (uint32_t)(e0 - s0);
        if (s0 & 1ull) {
            const uint4 r = philox4x32_10(plo, phi, c2, c3, keys);
            pack_step<ALGO, LOWT, MEAS, HIST>(k, c, r.z, r.w, table(r.w));
            ++plo; --left;
        }
        if (left >= 2u) {
            // Two pairs per trip, the blocks alternating between two register sets (no copies): while the
            // steps of one pair walk their dependent chain, the Philox block and the table entries of the
            // next pair -- independent of the chain -- fill the issue slots and hide behind its latency.
            uint4 rA = philox4x32_10(plo, phi, c2, c3, keys);
            uint4 a0 = table(rA.y), a1 = table(rA.w);
#if !WORM_PACK_UNROLL2
            for (; left >= 4u; left -= 2u) {
                ++plo;
                const uint4 rn = philox4x32_10(plo, phi, c2, c3, keys);
                const uint4 n0 = table(rn.y), n1 = table(rn.w);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.x, rA.y, a0);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.z, rA.w, a1);
                rA = rn; a0 = n0; a1 = n1;
            }
#endif
            for (; left >= 4u; left -= 4u) {
                const uint4 rB = philox4x32_10(plo + 1u, phi, c2, c3, keys);
                const uint4 b0 = table(rB.y), b1 = table(rB.w);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.x, rA.y, a0);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.z, rA.w, a1);
                plo += 2u;
                rA = philox4x32_10(plo, phi, c2, c3, keys);
                a0 = table(rA.y); a1 = table(rA.w);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rB.x, rB.y, b0);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rB.z, rB.w, b1);
            }
            if (left >= 2u) {
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.x, rA.y, a0);
                pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.z, rA.w, a1);
                left -= 2u;
                if (left) {
                    rA = philox4x32_10(plo + 1u, phi, c2, c3, keys);
                    a0 = table(rA.y);
                }
            }
            if (left) pack_step<ALGO, LOWT, MEAS, HIST>(k, c, rA.x, rA.y, a0);
        } else if (left) {
            const uint4 r = philox4x32_10(plo, phi, c2, c3, keys);
            pack_step<ALGO, LOWT, MEAS, HIST>(k, c, r.x, r.y, table(r.y));
        }
    };

    // first dump point at or after the first measured step of this launch
    uint64_t next_dump = ~0ull;
    if (a.dump_every) {
        const uint64_t mb = a.step_begin > a.therm ? a.step_begin : a.therm;
        next_dump = (mb + a.dump_every - 1) / a.dump_every * a.dump_every;
    }
    uint64_t s = a.step_begin;
    if (!has_col) s = a.step_end;                                  // lanes without a column only help with the copies
    while (s < a.step_end) {
        const bool meas = s >= a.therm;
        if (s == next_dump) {                                      // worm_ising_2d.cpp:113-124
            const bool want = live && (c.hs == c.tail);
            const bool store = want && nd < a.max_dumps;
            if (store && a.events) {
                int64_t *ev = a.events + ((int64_t)cl * a.max_dumps + nd) * 3;
                ev[0] = (int64_t)s; ev[1] = (int64_t)(c.Z + c.Zp); ev[2] = (int64_t)(c.SNb + c.SNbp);
            }
            unsigned dm = __ballot_sync(wmask, store && a.dumps != nullptr);
            while (dm) {
                const int src_lane = __ffs(dm) - 1;
                dm &= dm - 1;
                const long long chain = __shfl_sync(wmask, (long long)cl, src_lane);
                const int slot = __shfl_sync(wmask, nd, src_lane);
                uint8_t *dst = a.dumps + (chain * a.max_dumps + slot) * g.nb2;
                const uint32_t col = k.base + (uint32_t)(src_lane - lane) * 4u;
                for (int w = lane; w < g.W; w += wlanes) unpack_word<ALGO>(lds32(col + (uint32_t)w * k.pitch), dst, w);
            }
            if (want) ++nd;
            next_dump += a.dump_every;
        }
        uint64_t e = a.step_end;
        if (!meas && a.therm < e) e = a.therm;
        if (next_dump < e) e = next_dump;
        if (e - s > (uint64_t)g.fold) e = s + g.fold;
        if ((s ^ (e - 1)) >> 33) e = (s | 0x1ffffffffull) + 1ull;
        if (meas) run_range(std::true_type{}, s, e);
        else run_range(std::false_type{}, s, e);
        pack_fold(c);
        if constexpr (HIST) {
            if (c.run) pack_hist_flush<ALGO>(k, c);
        }
        s = e;
    }

    // ---- write back: all 32 lanes of a warp unpack one chain's column together (coalesced stores)
    __syncwarp();
    const bool writable_any = __any_sync(0xffffffffu, writable);
    int colsum = 0;
    if (writable_any) {
        for (int j = 0; j < wcols; ++j) {
            const bool wj = __shfl_sync(0xffffffffu, (int)writable, j) != 0;
            const long long chain = __shfl_sync(0xffffffffu, (long long)cl, j);
            if (!wj) continue;
            uint8_t *dst = a.bonds + chain * (long long)g.nb2;
            const uint32_t col = k.base + (uint32_t)(j - lane) * 4u;
            int sum = 0;
            for (int w = lane; w < g.W; w += 32) {
                const uint32_t word = lds32(col + (uint32_t)w * k.pitch);
                unpack_word<ALGO>(word, dst, w);
                if constexpr (ALGO == ALGO_TAYLOR)
                    sum += (int)((((word & 0x0f0f0f0fu) + ((word >> 4) & 0x0f0f0f0fu)) * 0x01010101u) >> 24);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == j) colsum = sum;
        }
    }
    if (writable) {
        // a nibble that carried into its neighbour (occupation 16) shows as a lost 16 in the column's sum
        if (ALGO == ALGO_TAYLOR && colsum != c.Nb) ac[ACC_STATUS] |= 2;
        a.head[cl] = (int32_t)(c.hs >> F::US);
        a.tail[cl] = (int32_t)(c.tail >> F::US);
        if constexpr (ALGO != ALGO_TAYLOR) { c.Sn = old3 + (c.SNb - old1); c.Sn2 = old4 + (c.SNb2 - old2); }
        const int64_t iters = measured_iters(a);
        const int64_t oldv[5] = {(int64_t)old0, (int64_t)old1, (int64_t)old2, (int64_t)old3, (int64_t)old4};
        add_group_acc(a, cl, oldv, c.Z, c.SNb, c.SNb2, c.Sn, c.Sn2, iters);
        ac[0] = (int64_t)c.Z; ac[1] = (int64_t)c.SNb; ac[2] = (int64_t)c.SNb2;
        ac[3] = (int64_t)c.Sn; ac[4] = (int64_t)c.Sn2;
        ac[5] += iters;
        if (a.n_dumps) a.n_dumps[cl] = nd;
    }
}

struct PackPlan {
    bool ok = false;
    PackGeo g{};
    size_t smem = 0;
    int64_t blocks = 0;
};

PackPlan pack_plan(int L, int algo, int64_t n_chains, int smem_optin, int sm_count)
{
    PackPlan p;
    if (L < 4 || (L & (L - 1)) || L > 2048) return p;
    const int N = L * L;
    const int bits = (algo == ALGO_TAYLOR) ? 4 : 1;
    const int W = 2 * N * bits / 32;
    if (W < 1 || (algo == ALGO_TAYLOR && (W & 1))) return p;
    const size_t per_chain = (size_t)W * 4;
    const size_t budget = (size_t)smem_optin > LUT_BYTES ? (size_t)smem_optin - LUT_BYTES : 0;
    int cmax = (int)(budget / per_chain);
    if (cmax < 1) return p;
    if (cmax > WORM_PACK_BIG) cmax = WORM_PACK_BIG;
    // a last partial warp costs the issue slots of a whole one: not worth it for a few chains
    if (cmax >= 32 && (cmax & 31) < 16) cmax &= ~31;
    // as many waves as the largest block needs, then the smallest block that still fits the batch in them:
    // the last wave is as full as the others
    const int64_t per_wave = (int64_t)cmax * sm_count;
    const int64_t waves = (n_chains + per_wave - 1) / per_wave;
    int64_t cb = (n_chains + waves * sm_count - 1) / (waves * sm_count);
    if (cb > cmax) cb = cmax;
    if (cb < 1) cb = 1;
    p.ok = true;
    p.g.L = L; p.g.N = N; p.g.nb2 = 2 * N; p.g.W = W;
    p.g.CB = (int)cb;
    p.g.TPB = ((int)cb + 31) & ~31;
    const int us = (algo == ALGO_TAYLOR) ? 3 : 1;
    int logN = 0;
    while ((1 << logN) < N) ++logN;
    p.g.ksh = 29 - logN - us;
    p.g.kmask = (uint32_t)(N - 1) << us;
    // 32-bit partial sums: a closed step adds at most 15 * 2N
    uint64_t fold = 0xffffffffull / (30ull * (uint64_t)N);
    if (fold > 16384) fold = 16384;
    fold &= ~1ull;
    if (fold < 2) fold = 2;
    p.g.fold = (uint32_t)fold;
    p.smem = LUT_BYTES + (size_t)W * p.g.CB * 4;
    p.smem = (p.smem + 15) & ~(size_t)15;
    if (p.smem > (size_t)smem_optin) { p.ok = false; return p; }
    p.blocks = (n_chains + cb - 1) / cb;
    // a multiplier coprime to the number of blocks, near the golden section of it
    auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
    int64_t mul = (int64_t)((double)p.blocks * 0.6180339887) | 1;
    while (mul > 1 && gcd(mul, p.blocks) != 1) mul -= 2;
    if (mul < 1 || p.blocks < 4) mul = 1;
    p.g.perm = (uint32_t)mul;
    return p;
}

template <int ALGO, bool LOWT, bool HIST, int MAXT>
cudaError_t launch_pack_t(const PhiloxArgs &a, const PackPlan &p, cudaStream_t st)
{
    auto kern = worm_pack_kernel<ALGO, LOWT, HIST, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)p.blocks, p.g.TPB, p.smem, st>>>(a, p.g);
    return cudaGetLastError();
}

template <int ALGO, bool LOWT, bool HIST>
cudaError_t launch_pack(const PhiloxArgs &a, const PackPlan &p, cudaStream_t st)
{
    if (p.g.TPB <= 256) return launch_pack_t<ALGO, LOWT, HIST, 256>(a, p, st);
    return launch_pack_t<ALGO, LOWT, HIST, WORM_PACK_BIG>(a, p, st);
}

template <int ALGO>
cudaError_t launch_pack_a(const PhiloxArgs &a, const PackPlan &p, bool lowt, bool hist, cudaStream_t st)
{
    if (lowt) return hist ? launch_pack<ALGO, true, true>(a, p, st) : launch_pack<ALGO, true, false>(a, p, st);
    return hist ? launch_pack<ALGO, false, true>(a, p, st) : launch_pack<ALGO, false, false>(a, p, st);
}

}  // namespace

int pack_chains_per_block(int L, int algo, int64_t n_chains, int smem_optin, int sm_count)
{
    const PackPlan p = pack_plan(L, algo, n_chains, smem_optin, sm_count);
    return p.ok ? p.g.CB : 0;
}

cudaError_t launch_philox_pack(const PhiloxArgs &a, bool lowt, int smem_optin, int sm_count, cudaStream_t st)
{
    const bool hist = (a.hist != nullptr && a.group_index != nullptr);
    const PackPlan p = pack_plan(a.L, a.algo, a.n_chains, smem_optin, sm_count);
    if (!p.ok) return cudaErrorInvalidValue;
    if (a.algo == ALGO_TAYLOR) return launch_pack_a<ALGO_TAYLOR>(a, p, lowt, hist, st);
    return launch_pack_a<ALGO_BINARY>(a, p, lowt, hist, st);
}

}  // namespace worm
