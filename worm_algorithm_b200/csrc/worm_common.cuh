// worm_common.cuh -- device helpers shared by the worm kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace worm {

// Per-chain parameters of the production (Philox) kernels, built on the host from (T, seed).
struct alignas(32) PhiloxChain {
    uint64_t kfix;   // ceil(2^32 * K),        K = 1.0/T                (Taylor increment test)
    uint64_t tfix;   // ceil(2^32 * (1.0/K))                            (Taylor decrement test)
    uint64_t hfix;   // ceil(2^32 * tanh(K))                            (binary worm)
    uint64_t seed;   // Philox key
};

// Per-chain parameters of the deterministic kernel: the doubles the reference computes at
// worm_ising_2d.cpp:49-50 and the 32-bit MT seed (worm_ising_2d.cpp:47, mt19937ar.c:70).
struct alignas(32) MtChain {
    double K;
    double inv_K;
    double T;
    uint32_t seed;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  10 rounds of {2 x 32x32->64 multiply, 2 x 3-input xor};
// the key schedule is a pure function of the key and is hoisted by the compiler.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The same function with the key schedule precomputed by the caller (kx[r] = k0 + r*0x9E3779B9,
// ky[r] = k1 + r*0xBB67AE85): a kernel that draws millions of blocks for one key keeps the twenty
// round keys in registers instead of re-adding them per block.
struct PhiloxKeys {
    uint32_t kx[10], ky[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1)
{
    PhiloxKeys k;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        k.kx[r] = k0 + (uint32_t)r * 0x9E3779B9u;
        k.ky[r] = k1 + (uint32_t)r * 0xBB67AE85u;
        asm volatile("" : "+r"(k.kx[r]), "+r"(k.ky[r]));     // opaque: keep them, do not rematerialise
    }
    return k;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k.kx[r];
        c2 = hi0 ^ c3 ^ k.ky[r];
        c1 = lo1;
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// floor(n / d) for n*d < 2^32 with magic = ceil(2^32 / d)  (exact in that range)
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, uint32_t magic) { return __umulhi(n, magic); }

}  // namespace worm
