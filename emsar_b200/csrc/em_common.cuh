// Device helpers shared by the EM kernels (em.cu: k_em_persistent<0..3>; em_psum.cu: k_em_psum).
#pragma once
#include "common.cuh"

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// a / b for finite a >= 0 and normal b > 0, within 1 ulp: hardware reciprocal seed (MUFU.RCP64H), two Newton steps, one
// residual correction - 8 instructions instead of the ~30 of the IEEE division sequence (one division per class and two per
// row every iteration: 8 % of all executed instructions, profiles/r1h). Tiny or huge divisors take the exact path.
__device__ __forceinline__ double fast_div(double a, double b)
{
    if (!(b > 1e-290 && b < 1e290)) return a / b;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// work queue of one chunk: items sorted by decreasing cost, warps take the next one (results do not depend on who computes)
__device__ __forceinline__ int next_item(int *counter, int lane)
{
    // one native shared-memory atomic by lane 0 (inline PTX: the compiler's warp-aggregated expansion of atomicAdd under a
    // divergent `if` was 15 % of all executed instructions, profiles/r1h)
    int t = 0;
    if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"((uint32_t)__cvta_generic_to_shared(counter)) : "memory");
    return __shfl_sync(0xffffffffu, t, 0);
}

// ---- tagged 16-byte slots: the exchange primitive of the fused sharded kernel ----------------------------------------
// (the LL idea of NCCL applied to fp64): the writer splits the double into two 32-bit halves and stores each together with
// the tag of the iteration in ONE aligned 8-byte word; the reader polls the two words until both carry the tag it expects.
// A value and its "ready" signal therefore travel in the same store: no system-scope fence, no barrier, and no assumption
// about the order in which different stores cross NVLink.
constexpr unsigned LL_PATIENCE = 1u << 22;      // polls (~ seconds) before a wait gives up and raises the abort flag
__device__ __forceinline__ void ll_store(unsigned char *slot, double v, unsigned tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v), t = (unsigned long long)tag << 32;
    const unsigned long long w0 = (b & 0xffffffffULL) | t, w1 = (b >> 32) | t;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ double ll_load(const unsigned char *slot, unsigned tag, int *abort_flag)
{
    unsigned long long w0, w1;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
        if ((unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag) break;
        if ((++spins & 1023u) == 0 && (*((volatile int *)abort_flag) != 0 || spins >= LL_PATIENCE)) { *((volatile int *)abort_flag) = 1; break; }
    }
    return __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
}

