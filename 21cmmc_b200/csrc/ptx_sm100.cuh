// sm_100a building blocks shared by the two-pass kernels (ps21_fused.cu, ps21_g2.cu): mbarrier + bulk asynchronous
// copies (TMA, 1-D), named barriers, L2 eviction policies, and the real-FFT untangle step.  The emulated test build
// (PS21_EMU) models a bulk copy as a synchronous memcpy.
#pragma once
#include "ps21_internal.cuh"

namespace ps21 {

// 32-bit shared-memory addressing: the generic kernels (ps21_g2.cu) form their addresses as 32-bit byte offsets in the
// shared window and access them with explicit ld.shared / st.shared -- pointer arithmetic on 64-bit generic pointers
// cost three integer instructions per access there (ncu: half of all instructions were address arithmetic).
#ifndef PS21_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)(reinterpret_cast<const char*>(p) - ps21emu::g_dyn_smem);
}
__device__ __forceinline__ float2 lds64(uint32_t a) { return *reinterpret_cast<const float2*>(ps21emu::g_dyn_smem + a); }
__device__ __forceinline__ void sts64(uint32_t a, float2 v) { *reinterpret_cast<float2*>(ps21emu::g_dyn_smem + a) = v; }
#endif

#ifndef PS21_EMU
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// barrier over `nthreads` threads of the CTA (named barrier `id`, 1..15)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// non-blocking arrival at named barrier `id` (the warps that call named_bar_sync on it wait for `nthreads` arrivals)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// L2 eviction policies: streamed data must not push the (re-used) bin-index table out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// contiguous bytes -> L2 only (no shared-memory destination, no completion to wait for); the range is widened to
// 16-byte boundaries, which the instruction requires
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(src);
    const unsigned long long a0 = a & ~15ull;
    const uint32_t n = (uint32_t)(((a + bytes + 15ull) & ~15ull) - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

#else
// emulated build (tests only): the bulk copy is a synchronous memcpy; the mbarrier is modelled as
// {pending bytes, arrived flag, phase bit} packed in its 8 bytes; an L2 prefetch is nothing
__device__ __forceinline__ void bulk_prefetch_l2(const void*, uint32_t) {}
struct EmuBar {
    int pending;
    unsigned short arrived, phase;
};
__device__ __forceinline__ void emu_bar_check(EmuBar* b) {
    if (b->arrived && b->pending == 0) {
        b->arrived = 0;
        b->phase ^= 1;
    }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->pending += (int)bytes;
    b->arrived = 1;
    emu_bar_check(b);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    while (b->phase == parity) ps21emu::spin_yield();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst, src, bytes);
    EmuBar* b = reinterpret_cast<EmuBar*>(bar);
    b->pending -= (int)bytes;
    emu_bar_check(b);
}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void named_bar_sync(int, int) { __syncthreads(); }  // stronger, same result
__device__ __forceinline__ void named_bar_arrive(int, int) { __syncthreads(); }
__device__ __forceinline__ uint64_t l2_policy_evict_first() { return 0; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 0; }
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t) {
    bulk_g2s(dst, src, bytes, bar);
}

#endif

// untangle of the (even,odd)-packed real FFT for the column pair (p, N-p).
// With E = (C_p + conj C_{N-p})/2 and O = (C_p - conj C_{N-p})/(2i):  A_p = E + w^p O,  A_{N-p} = conj(E - w^p O).
// `wu` is the table entry of  w_256^p * (-i/2), so that  w^p O = wu * (C_p - conj C_{N-p}).
__device__ __forceinline__ void untangle(float2 ck, float2 cn, float4 wu, float2& a1, float2& a2) {
#ifdef PS21_PACKED_F32X2
    const float2 e2 = __ffma2_rn(cn, make_float2(1.f, -1.f), ck);   // 2E
    const float2 d = __ffma2_rn(cn, make_float2(-1.f, 1.f), ck);    // C_p - conj C_{N-p}
    const float2 t = cmul_tw(d, wu);
    a1 = __ffma2_rn(e2, make_float2(0.5f, 0.5f), t);
    a2 = __ffma2_rn(e2, make_float2(0.5f, -0.5f), __fmul2_rn(t, make_float2(-1.f, 1.f)));
#else
    const float2 e2 = make_float2(ck.x + cn.x, ck.y - cn.y);
    const float2 d = make_float2(ck.x - cn.x, ck.y + cn.y);
    const float2 t = cmul_tw(d, wu);
    a1 = make_float2(0.5f * e2.x + t.x, 0.5f * e2.y + t.y);
    a2 = make_float2(0.5f * e2.x - t.x, -(0.5f * e2.y - t.y));
#endif
}


}  // namespace ps21
