/*
 * worm_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement; never on the product path).
 *
 * Plain-C restatement of the reference's worm-algorithm hot path, written from the
 * behaviour of /root/reference/worm_algorithm/src/worm_ising_2d.cpp (cited per function
 * below as "ref:LINE") and of the published MT19937 and Philox4x32-10 algorithms.
 *
 * Parity status: PINNED.  worm_oracle_mt() is checked in tests/test_oracle.py against the
 * golden vectors in tests/golden/ref_*.json, which oracle/make_golden.py produced by running
 * the UNMODIFIED reference executable (oracle/_ref/worm_ising_2d, built by oracle/Makefile),
 * and -- when that executable is present -- against fresh runs of it.  The reference itself
 * ships no tests or golden vectors.
 *
 * Three entry points:
 *   worm_oracle_mt      deterministic mode: the reference's MT19937 stream, double-precision
 *                       acceptance compares exactly as the C++ writes them          (ref:111-157)
 *   worm_oracle_philox  production mode, restated on the CPU so the CUDA kernels can be checked
 *                       bit-for-bit: algo 0 = the reference's integer-occupation (Taylor) worm,
 *                       algo 1 = binary-bond tanh(K) worm; counter-based Philox4x32-10 stream.
 *   worm_oracle_thresholds   the fixed-point acceptance constants both sides use.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura, ACM TOMACS 8(1), 1998), 2002 initialisation.
 * Behaviour pinned to ref mt19937ar.c:68-81 (init_genrand) and :113-148 (genrand_int32).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t s[624];
    int pos;
} mt_state;

static void mt_seed(mt_state *g, uint32_t seed)
{
    g->s[0] = seed;
    for (int i = 1; i < 624; i++) {
        uint32_t p = g->s[i - 1];
        g->s[i] = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
    }
    g->pos = 624;
}

static void mt_refill(mt_state *g)
{
    for (int i = 0; i < 624; i++) {
        uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
        uint32_t v = g->s[(i + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        g->s[i] = v;
    }
    g->pos = 0;
}

static uint32_t mt_u32(mt_state *g)
{
    if (g->pos >= 624) mt_refill(g);
    uint32_t y = g->s[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* ref mt19937ar.c:185-189: genrand() = genrand_int32() * 2^-32 */
static double mt_real(mt_state *g) { return mt_u32(g) * (1.0 / 4294967296.0); }

/* first n raw words of the stream, for the RNG unit test */
void worm_oracle_mt_stream(uint32_t seed, int64_t n, uint32_t *out)
{
    mt_state g;
    mt_seed(&g, seed);
    for (int64_t i = 0; i < n; i++) out[i] = mt_u32(&g);
}

/* ------------------------------------------------------------------------------------------
 * Lattice helpers.  site = y*L + x; neighbour directions 0:+x 1:+y 2:-x 3:-y (ref:88-95).
 * ---------------------------------------------------------------------------------------- */
static int nbr_site(int site, int d, int L)
{
    int N = L * L;
    switch (d) {
    case 0: return (site / L) * L + (site + 1 + N) % L;
    case 1: return (site + L) % N;
    case 2: return (site / L) * L + (site - 1 + N) % L;
    default: return (site - L + N) % N;
    }
}

/* ref:214-244 bond_number: the branch order matters for L == 2. */
static int bond_index(int i1, int i2, int L)
{
    int x1 = i1 % L, y1 = i1 / L, x2 = i2 % L, y2 = i2 / L;
    if (y1 == y2) {
        if (x2 == x1 + 1) return 2 * i1;
        if (x1 == x2 + 1) return 2 * i2;
        if (x1 == L - 1) return 2 * i1;
        if (x2 == L - 1) return 2 * i2;
    } else if (x1 == x2) {
        if (y2 == y1 + 1) return 2 * i1 + 1;
        if (y1 == y2 + 1) return 2 * i2 + 1;
        if (y1 == L - 1) return 2 * i1 + 1;
        if (y2 == L - 1) return 2 * i2 + 1;
    }
    return -1; /* the reference prints "ERROR!" and indexes out of range here (ref:242-243) */
}

/* bond map rows "bond x1 y1 x2 y2", 4N of them, in the order ref:169-181 writes them */
void worm_oracle_bond_map(int L, int64_t *rows /* [4N][5] */)
{
    int N = L * L;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < 4; j++) {
            int n2 = nbr_site(i, j, L);
            int64_t *r = rows + ((int64_t)i * 4 + j) * 5;
            r[0] = bond_index(i, n2, L);
            r[1] = i % L; r[2] = i / L; r[3] = n2 % L; r[4] = n2 / L;
        }
}

/* ------------------------------------------------------------------------------------------
 * Deterministic mode: the reference main loop, ref:47-57 (setup) and ref:111-157 (loop).
 *
 *   events[e] = {step_num, Z, Nb_tot} at every point the reference writes an observables row
 *               (ref:113-119: closed, step > 0.1*num_steps, step % 50 == 0), values BEFORE the
 *               kick of that iteration; when bond_flag != 0 the bond array is dumped there too
 *               (ref:120-124) and once more after the loop (ref:160-167).
 *   trace[k]  = {step_num, Nb} at the k-th closed iteration (the N_b trace; the reference
 *               folds it into Nb_tot, ref:132).
 *   out[0..3] = Z, Nb_tot (negative running sum, 64-bit here; `int` in the reference), step_num, Nb.
 * Returns 0, or -1 on a bad edge, -2 on bad arguments.
 * ---------------------------------------------------------------------------------------- */
int worm_oracle_mt(int L, double T, uint64_t num_steps, double seed, int bond_flag,
                   int64_t *out, int32_t *final_bonds,
                   int64_t max_events, int64_t *n_events, int64_t *events, int32_t *dumps,
                   int64_t max_trace, int64_t *n_trace, int64_t *trace)
{
    if (L < 2) return -2;
    const int N = L * L;
    mt_state g;
    mt_seed(&g, (uint32_t)(unsigned long)seed);          /* ref:47; mt19937ar.c:70 keeps 32 bits */
    const double K = 1.0 / T;                              /* ref:49 */
    const double inv_K = 1.0 / K;                          /* ref:50 */
    int tail = 0, head = 0, Nb = 0;
    int64_t Nb_tot = 0;
    uint64_t step = 0, Z = 0;
    const uint64_t therm = (uint64_t)(0.1 * (double)num_steps);   /* ref:55 */
    int32_t *bonds = (int32_t *)calloc((size_t)2 * N, sizeof(int32_t));
    int64_t ne = 0, nt = 0;
    int rc = 0;
    for (;;) {
        if (tail == head) {                                                    /* ref:112 */
            if (step > therm && step % 50 == 0) {                              /* ref:113-114 */
                if (ne < max_events) {
                    if (events) { events[3 * ne] = (int64_t)step; events[3 * ne + 1] = (int64_t)Z; events[3 * ne + 2] = Nb_tot; }
                    if (dumps && bond_flag) memcpy(dumps + (size_t)ne * 2 * N, bonds, sizeof(int32_t) * 2 * N);
                }
                ne++;
            }
            if (nt < max_trace && trace) { trace[2 * nt] = (int64_t)step; trace[2 * nt + 1] = Nb; }
            nt++;
            tail = (int)floor(mt_real(&g) * N);                                /* ref:128 */
            head = tail;
            Z += 1;                                                            /* ref:130 */
            Nb_tot -= Nb;                                                      /* ref:132 */
            if (step >= num_steps) break;                                      /* ref:133-134 */
        }
        int new_head = nbr_site(head, (int)floor(mt_real(&g) * 4), L);         /* ref:137 */
        int ib = bond_index(head, new_head, L);                                /* ref:138 */
        if (ib < 0) { rc = -1; break; }
        int nb = bonds[ib];                                                    /* ref:139 */
        int delta;
        double P;
        if (mt_real(&g) < 0.5) { delta = 1;  P = K / (nb + 1.0); }             /* ref:141-144 */
        else                   { delta = -1; P = nb * inv_K; }                 /* ref:146-149 */
        if (mt_real(&g) < P) {                                                 /* ref:151 */
            bonds[ib] += delta;
            Nb += delta;
            head = new_head;
        }
        ++step;                                                                /* ref:156 */
    }
    if (final_bonds) memcpy(final_bonds, bonds, sizeof(int32_t) * 2 * N);
    out[0] = (int64_t)Z; out[1] = Nb_tot; out[2] = (int64_t)step; out[3] = Nb;
    if (n_events) *n_events = ne;
    if (n_trace) *n_trace = nt;
    free(bonds);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Integer-exact acceptance thresholds: genrand() < P  <=>  u32 < thr(P),
 * thr(P) = clamp(ceil(P * 2^32), 0, 2^32).  P*2^32 is an exact scaling of the double the
 * reference computes, so the table reproduces ref:143,148,151 for every 32-bit draw.
 * thr_inc[nb] = thr(K/(nb+1.0)), thr_dec[nb] = thr(nb*inv_K), nb = 0..255.
 * ---------------------------------------------------------------------------------------- */
static uint64_t thr_of(double P)
{
    double s = ceil(ldexp(P, 32));
    if (!(s > 0.0)) return 0;
    if (s >= 4294967296.0) return 4294967296ull;
    return (uint64_t)s;
}

void worm_oracle_mt_tables(double T, uint64_t *thr_inc /*[256]*/, uint64_t *thr_dec /*[256]*/)
{
    const double K = 1.0 / T, inv_K = 1.0 / K;
    for (int nb = 0; nb < 256; nb++) {
        thr_inc[nb] = thr_of(K / (nb + 1.0));
        thr_dec[nb] = thr_of(nb * inv_K);
    }
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Known-answer vectors are checked in
 * tests/test_oracle.py.
 * ---------------------------------------------------------------------------------------- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void worm_oracle_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    philox4x32_10(ctr, key, out);
}

/* Fixed-point constants of the production acceptance tests:
 *   kfix = ceil(K * 2^32), K = 1.0/T          (increment:  (nb+1) * u < kfix   <=> u/2^32 < K/(nb+1))
 *   tfix = ceil((1.0/K) * 2^32)               (decrement:  u < nb * tfix       <=> u/2^32 < nb/K)
 *   hfix = ceil(tanh(K) * 2^32)               (binary worm: empty bond accepted iff u < hfix)
 * They differ from the reference's double-rounded K/(nb+1.0) by < 2^-32 relative. */
void worm_oracle_thresholds(double T, uint64_t *kfix, uint64_t *tfix, uint64_t *hfix)
{
    const double K = 1.0 / T, inv_K = 1.0 / K;
    *kfix = (uint64_t)ceil(ldexp(K, 32));
    *tfix = (uint64_t)ceil(ldexp(inv_K, 32));
    *hfix = (uint64_t)ceil(ldexp(tanh(K), 32));
}

/* ------------------------------------------------------------------------------------------
 * Production mode.  Random stream (version 2): ONE Philox4x32-10 call serves TWO worm steps.
 *   pair p = step >> 1, counter = {p_lo, p_hi, chain_lo, chain_hi}, key = {seed_lo, seed_hi},
 *   output (x, y, z, w); an even step uses (a, b) = (x, y), an odd step (a, b) = (z, w):
 *     a                      acceptance draw, all 32 bits                        (ref:151)
 *     b >> 30                direction 0..3                                      (ref:137)
 *     (b >> 29) & 1          0 = increment proposal, 1 = decrement proposal      (ref:141)
 *     ((b << 3) * N) >> 32   kick site, from the low 29 bits of b                (ref:128)
 *   (the reference spends one 32-bit draw on each of the four; the three small decisions use
 *    disjoint bit fields of one word here.  For power-of-two N the kick is exactly uniform,
 *    otherwise its bias is below N * 2^-29.)
 *
 * The loop runs steps [step_begin, step_begin + n_steps) of one chain, resuming from
 * (bonds, head, tail).  Iterations with step >= therm are measured:
 *   closed (head == tail):  acc[0] Z += 1, acc[1] += Nb, acc[2] += Nb^2, acc[3] += n, acc[4] += n^2
 *       (Nb = sum of occupations, the reference's estimator ref:132,185; n = number of odd-parity
 *        bonds, what count_bonds.py:110-123 counts on the images).  algo 1: Nb == n == occupied bonds.
 *   every iteration:        hist[dy*L + dx] += 1, (dx,dy) = head - tail mod L, taken BEFORE the kick
 *       (so hist[0] == Z).
 *   closed && dump_every > 0 && step % dump_every == 0: event {step, Z, sum Nb} (values before this
 *       iteration's update, like ref:115-119) and a dump of the occupations (ref:120-124).
 * acc[5] counts measured iterations.  Returns 0 or -2.
 * ---------------------------------------------------------------------------------------- */
int worm_oracle_philox_run(int L, int algo, uint64_t chain_id, uint64_t seed,
                           uint64_t kfix, uint64_t tfix, uint64_t hfix,
                           uint64_t step_begin, uint64_t n_steps, uint64_t therm,
                           uint8_t *bonds /* [2N] in/out */, int32_t *head_io, int32_t *tail_io,
                           int64_t *acc /* [8] in/out */, int64_t *hist /* [N] or NULL, in/out */,
                           uint64_t dump_every, int64_t max_dumps, int64_t *n_dumps_io,
                           int64_t *events /* [max_dumps][3] */, uint8_t *dumps /* [max_dumps][2N] or NULL */)
{
    if (L < 3 || (algo != 0 && algo != 1)) return -2;
    const int N = L * L;
    int head = *head_io, tail = *tail_io;
    int64_t Nb = 0, npar = 0;
    for (int b = 0; b < 2 * N; b++) { Nb += bonds[b]; npar += bonds[b] & 1; }
    int64_t nd = n_dumps_io ? *n_dumps_io : 0;
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (uint64_t s = step_begin; s < step_begin + n_steps; s++) {
        const uint64_t pair = s >> 1;
        const uint32_t ctr[4] = { (uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)chain_id, (uint32_t)(chain_id >> 32) };
        uint32_t r[4];
        philox4x32_10(ctr, key, r);
        const uint32_t ra = (s & 1) ? r[2] : r[0];      /* acceptance draw */
        const uint32_t rb = (s & 1) ? r[3] : r[1];      /* direction | coin | kick bits */
        const int closed = (head == tail);
        if (s >= therm) {
            if (hist) {
                int dx = head % L - tail % L; if (dx < 0) dx += L;
                int dy = head / L - tail / L; if (dy < 0) dy += L;
                hist[dy * L + dx] += 1;
            }
            if (closed) {
                if (dump_every && s % dump_every == 0) {
                    if (nd < max_dumps) {
                        if (events) { events[3 * nd] = (int64_t)s; events[3 * nd + 1] = acc[0]; events[3 * nd + 2] = acc[1]; }
                        if (dumps) memcpy(dumps + (size_t)nd * 2 * N, bonds, (size_t)2 * N);
                    }
                    nd++;
                }
                acc[0] += 1; acc[1] += Nb; acc[2] += Nb * Nb; acc[3] += npar; acc[4] += npar * npar;
            }
            acc[5] += 1;
        }
        if (closed) head = tail = (int)(((uint64_t)(uint32_t)(rb << 3) * (uint64_t)N) >> 32);
        const int d = (int)(rb >> 30);
        const int x = head % L, y = head / L;
        int nh, ib;
        switch (d) {                       /* closed form of ref:88-95 + ref:214-244, valid for L >= 3 */
        case 0: nh = y * L + (x + 1) % L;           ib = 2 * head;     break;
        case 1: nh = ((y + 1) % L) * L + x;         ib = 2 * head + 1; break;
        case 2: nh = y * L + (x + L - 1) % L;       ib = 2 * nh;       break;
        default: nh = ((y + L - 1) % L) * L + x;    ib = 2 * nh + 1;   break;
        }
        const uint32_t nb = bonds[ib];
        int accept, delta;
        if (algo == 0) {
            if (((rb >> 29) & 1u) == 0) { delta = 1;  accept = (nb < 255) && ((uint64_t)(nb + 1) * ra < kfix); }
            else                        { delta = -1; accept = ((uint64_t)ra < (uint64_t)nb * tfix); }
        } else {
            if (nb) { delta = -1; accept = 1; }
            else    { delta = 1;  accept = ((uint64_t)ra < hfix); }
        }
        if (accept) {
            const uint32_t nn = nb + delta;
            bonds[ib] = (uint8_t)nn;
            Nb += delta;
            npar += (nn & 1) ? 1 : -1;
            head = nh;
        }
    }
    *head_io = head; *tail_io = tail;
    if (n_dumps_io) *n_dumps_io = nd;
    return 0;
}
