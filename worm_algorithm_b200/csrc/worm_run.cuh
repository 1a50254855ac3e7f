// worm_run.cuh -- pieces shared by the two production kernels (worm_wide.cu, worm_lat.cu).
#pragma once
#include "worm_common.cuh"
#include "worm_kernels.h"

namespace worm {

constexpr int ALGO_TAYLOR = 0;
constexpr int ALGO_BINARY = 1;

// the two words of step s out of the Philox block of its pair
__device__ __forceinline__ void step_words(const uint4 r, bool odd, uint32_t &a, uint32_t &b)
{
    a = odd ? r.z : r.x;
    b = odd ? r.w : r.y;
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
// be its FIRST step.  The next dump multiple is carried along (one 64-bit modulo per launch: on a lone
// in-order warp a modulo per segment cost ~500 cycles of the 2900 a 50-step segment takes).
struct Segments {
    uint64_t next_dump;                  // smallest multiple of dump_every >= the current step (measured steps only)
    __device__ __forceinline__ explicit Segments(const PhiloxArgs &a)
    {
        next_dump = ~0ull;
        if (a.dump_every) {
            const uint64_t s0 = a.step_begin > a.therm ? a.step_begin : a.therm;
            const uint64_t rem = s0 % a.dump_every;
            next_dump = rem ? s0 + (a.dump_every - rem) : s0;
        }
    }
    // returns seg_end, sets meas / dump_first.  cut_dumps = false: segments end at MEAS changes only and the
    // caller handles the dump multiples inside its rounds (next_dump is its to advance).
    __device__ __forceinline__ uint64_t next(const PhiloxArgs &a, uint64_t s, bool &meas, bool &dump_first, bool cut_dumps = true)
    {
        meas = (s >= a.therm);
        uint64_t seg_end = a.step_end;
        if (!meas && a.therm < seg_end) seg_end = a.therm;
        dump_first = false;
        if (meas && a.dump_every && cut_dumps) {
            dump_first = (s == next_dump);
            if (dump_first) next_dump += a.dump_every;
            if (next_dump < seg_end) seg_end = next_dump;
        }
        return seg_end;
    }
};

}  // namespace worm
