// sspsd_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Replaces, on the device, the per-stage hot loop of the reference (src/psd.rs:196-269):
//   Detrend::apply (psd.rs:75-113) -> rustfft Fft::process (psd.rs:213) -> |X|^2 accumulate
//   (psd.rs:228-233) -> idsp HBF_DEC_CASCADE decimate-by-8 (psd.rs:246-253).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sspsd {

// ---------------------------------------------------------------------------------------------
// A stage's input stream as seen by a kernel: logical sample index g (0 = first sample the stage
// ever received) lives in `carry` for g < split (history + pending samples kept from earlier
// batches, zeros for g < 0) and in `fresh` for g >= split (this batch).  carry_start and split
// are multiples of 4 and both bases are 16-byte aligned, so aligned float4 groups never straddle.
// ---------------------------------------------------------------------------------------------
struct StreamSrc {
    const float* carry;
    const float* fresh;
    long long carry_start;
    long long split;
};

__device__ __forceinline__ float4 ld_stream4(const StreamSrc& s, long long g)
{
    if (g >= s.split)
        return __ldg(reinterpret_cast<const float4*>(s.fresh + (g - s.split)));
    if (g >= s.carry_start)
        return *reinterpret_cast<const float4*>(s.carry + (g - s.carry_start));
    return make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float ld_stream1(const StreamSrc& s, long long g)
{
    if (g >= s.split)
        return __ldg(s.fresh + (g - s.split));
    if (g >= s.carry_start)
        return s.carry[g - s.carry_start];
    return 0.f;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion reported to an mbarrier in bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// issue the copies of stream samples [g, g + n) into dst (n multiple of 4, g multiple of 4)
__device__ __forceinline__ void ring_issue(const StreamSrc& s, long long g, int n, float* dst, uint64_t* bar)
{
    mbar_expect_tx(bar, (uint32_t)n * 4u);
    long long nc = s.split - g;  // samples that live in the carry
    if (nc > n) nc = n;
    if (nc > 0) {
        bulk_g2s(dst, s.carry + (g - s.carry_start), (uint32_t)nc * 4u, bar);
    } else {
        nc = 0;
    }
    if (nc < n) bulk_g2s(dst + nc, s.fresh + (g + nc - s.split), (uint32_t)(n - nc) * 4u, bar);
}

// ---------------------------------------------------------------------------------------------
// FFT plan: an N-point real segment is transformed as an M = N/2 point complex FFT of
// z[n] = x[2n] + i x[2n+1] followed by the real-input split.  The complex FFT is an in-place
// decimation-in-frequency transform in shared memory; every thread owns 8 points per pass.
// Radices: as many radix-8 passes as fit, one radix-2 or radix-4 pass for the remainder, and a
// final radix-4 pass on 4 contiguous points (which is where the split pairs k with M-k).
// ---------------------------------------------------------------------------------------------
template <int LOG2N>
struct Plan {
    static constexpr int N = 1 << LOG2N;
    static constexpr int LOG2M = LOG2N - 1;
    static constexpr int M = 1 << LOG2M;
    static constexpr int TPS = M / 8;                      // threads per segment
    static constexpr int NT = TPS > 256 ? TPS : 256;       // threads per CTA
    static constexpr int G = NT / TPS;                     // segments in flight per CTA
    static constexpr int REM = LOG2M - 2;
    static constexpr int A = REM / 3;                      // radix-8 passes
    static constexpr int EXTRA = REM % 3;                  // log2 radix of the odd pass (0 = none)
    static constexpr int P = A + (EXTRA ? 1 : 0) + 1;      // total passes
    static constexpr int K = M / 4;                        // butterflies of the last pass
    static constexpr int WS = M + 4 * (M / 32);            // padded floats per re/im plane
    static_assert(LOG2N >= 6 && LOG2N <= 13, "supported FFT sizes: 64..8192");
    static_assert(A >= 1, "first pass must be radix 8");

    __host__ __device__ static constexpr int log2radix(int p) { return p < A ? 3 : (p == P - 1 ? 2 : EXTRA); }
    // log2 of the element stride inside pass p's butterflies (= size of the sub-transforms left)
    __host__ __device__ static constexpr int log2stride(int p)
    {
        int s = LOG2M;
        for (int q = 0; q <= p; ++q) s -= log2radix(q);
        return s;
    }
    // log2 of the weight of pass p's output digit in the frequency index
    __host__ __device__ static constexpr int log2weight(int p)
    {
        int s = 0;
        for (int q = 0; q < p; ++q) s += log2radix(q);
        return s;
    }
};

// padded shared-memory position of complex element i (4 floats of padding per 32 elements)
__device__ __forceinline__ int ws_pos(int i) { return i + ((i >> 5) << 2); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-place forward DFTs, v[k] = sum_t v[t] exp(-2 pi i t k / R)
__device__ __forceinline__ void dft2(float2& a, float2& b)
{
    float2 t = a;
    a = make_float2(t.x + b.x, t.y + b.y);
    b = make_float2(t.x - b.x, t.y - b.y);
}

__device__ __forceinline__ void dft4(float2& c0, float2& c1, float2& c2, float2& c3)
{
    float2 e0 = make_float2(c0.x + c2.x, c0.y + c2.y);
    float2 e1 = make_float2(c0.x - c2.x, c0.y - c2.y);
    float2 f0 = make_float2(c1.x + c3.x, c1.y + c3.y);
    float2 f1 = make_float2(c1.y - c3.y, -(c1.x - c3.x)); // -i (c1 - c3)
    c0 = make_float2(e0.x + f0.x, e0.y + f0.y);
    c2 = make_float2(e0.x - f0.x, e0.y - f0.y);
    c1 = make_float2(e1.x + f1.x, e1.y + f1.y);
    c3 = make_float2(e1.x - f1.x, e1.y - f1.y);
}

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R])
{
    if constexpr (R == 2) {
        dft2(v[0], v[1]);
    } else if constexpr (R == 4) {
        dft4(v[0], v[1], v[2], v[3]);
    } else {
        static_assert(R == 8, "radix");
        constexpr float h = 0.70710678118654752440f;
        float2 a0 = make_float2(v[0].x + v[4].x, v[0].y + v[4].y);
        float2 a1 = make_float2(v[1].x + v[5].x, v[1].y + v[5].y);
        float2 a2 = make_float2(v[2].x + v[6].x, v[2].y + v[6].y);
        float2 a3 = make_float2(v[3].x + v[7].x, v[3].y + v[7].y);
        float2 b0 = make_float2(v[0].x - v[4].x, v[0].y - v[4].y);
        float2 d1 = make_float2(v[1].x - v[5].x, v[1].y - v[5].y);
        float2 d2 = make_float2(v[2].x - v[6].x, v[2].y - v[6].y);
        float2 d3 = make_float2(v[3].x - v[7].x, v[3].y - v[7].y);
        float2 b1 = make_float2(h * (d1.x + d1.y), h * (d1.y - d1.x));  // * (1 - i)/sqrt2
        float2 b2 = make_float2(d2.y, -d2.x);                            // * -i
        float2 b3 = make_float2(h * (d3.y - d3.x), -h * (d3.x + d3.y)); // * (-1 - i)/sqrt2
        dft4(a0, a1, a2, a3);
        dft4(b0, b1, b2, b3);
        v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
        v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
    }
}

}  // namespace sspsd
