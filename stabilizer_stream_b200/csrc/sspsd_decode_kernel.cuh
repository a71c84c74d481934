// sspsd_decode_kernel.cuh -- K1: batched stabilizer frame decode + loss accounting.
//
// Replaces, for a batch of equally sized frames, the per-frame host loop of the reference:
//   Frame::from_bytes / Header::parse  (src/de/frame.rs:25-60)
//   Loss::update                       (src/loss.rs:11-26)
//   Payload::traces                    (src/de/data.rs:28-82, 97-139, 154-163, 178-211)
// Integer/byte work, HBM bound: headers are validated by one thread per frame, loss is a block
// reduced sum of wrapping u32 gaps between neighbouring headers (bit exact, order independent), and
// the AdcDac payload (the high-rate format) is scanned as a flat array of aligned 128-bit loads:
// each 8-byte word is four i16 samples of one channel and becomes one float4 store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sspsd.h"

namespace sspsd {

struct DecodeResult {
    unsigned long long first_bad;  // index of the first malformed frame (n_frames if none)
    unsigned long long received;   // sum of batches over frames < first_bad
    unsigned long long dropped;    // sum of wrapping gaps over frames < first_bad
    unsigned int status;           // error code of frame first_bad
    unsigned int last_seq_end;     // seq + batches of frame first_bad-1
    unsigned int format;           // format of frame 0
    unsigned int batches;          // batches of frame 0
};

struct DecodeParams {
    const uint8_t* frames;
    unsigned long long n_frames;
    unsigned long long frame_len;
    unsigned long long frame_stride;
    unsigned int* status;  // per frame: 0 or error code
    DecodeResult* res;
    unsigned int prev_seq;  // loss.seq
    unsigned int has_prev;
};

__host__ __device__ __forceinline__ unsigned int batch_bytes(unsigned int format)
{
    // AdcDac [[[[u8;2];8];4]] = 64, Fls [[[u8;4];7];2] = 56, ThermostatEem [[u8;4];20] = 80, Mpll [[u8;4];6] = 24
    return format == 1 ? 64u : format == 2 ? 56u : format == 3 ? 80u : 24u;
}

__device__ __forceinline__ unsigned int rd_u32(const uint8_t* p)
{
    return (unsigned int)p[0] | ((unsigned int)p[1] << 8) | ((unsigned int)p[2] << 16) | ((unsigned int)p[3] << 24);
}

// status of one frame; fmt0 = format the batch must have (0: take this frame's)
__host__ __device__ __forceinline__ unsigned int frame_status(const uint8_t* f, unsigned long long len, unsigned int fmt0)
{
    if (len < SSPSD_HEADER_SIZE) return SSPSD_ESHORT;            // frame.rs:50
    if (f[0] != 0x7b || f[1] != 0x05) return SSPSD_EHEADER;      // frame.rs:26-28
    unsigned int fmt = f[2];
    if (fmt < 1 || fmt > 4) return SSPSD_EFORMAT;                // frame.rs:29
    if (fmt0 && fmt != fmt0) return SSPSD_EFORMAT;               // one format per batched call
    unsigned long long dl = len - SSPSD_HEADER_SIZE;
    unsigned int bs = batch_bytes(fmt);
    if (dl % bs) return SSPSD_ESIZE;                             // bytemuck::try_cast_slice
    if (dl / bs != f[3]) return SSPSD_EBATCHES;                  // assert_eq!(data.len(), batches)
    return SSPSD_OK;
}

__global__ void frame_scan_kernel(const DecodeParams p)
{
    unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.n_frames) return;
    const uint8_t* fr = p.frames + f * p.frame_stride;
    unsigned int fmt0 = p.frames[2];
    unsigned int st = frame_status(fr, p.frame_len, (fmt0 >= 1 && fmt0 <= 4) ? fmt0 : 0);
    p.status[f] = st;
    if (st != SSPSD_OK) atomicMin(&p.res->first_bad, f);
}

// loss over frames < first_bad (Loss::update applied in order == sum of neighbour gaps)
__global__ void loss_kernel(const DecodeParams p)
{
    const unsigned long long nb = p.res->first_bad;
    unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long rec = 0, drop = 0;
    if (f < nb) {
        const uint8_t* fr = p.frames + f * p.frame_stride;
        unsigned int seq = rd_u32(fr + 4);
        unsigned int bat = fr[3];
        rec = bat;
        if (f > 0) {
            const uint8_t* pr = fr - p.frame_stride;
            unsigned int expect = rd_u32(pr + 4) + pr[3];  // wrapping_add, loss.rs:25
            drop = (unsigned int)(seq - expect);            // wrapping_sub, loss.rs:14
        } else if (p.has_prev) {
            drop = (unsigned int)(seq - p.prev_seq);
        }
        if (f == nb - 1) p.res->last_seq_end = seq + bat;
        if (f == 0) {
            p.res->format = fr[2];
            p.res->batches = bat;
        }
    }
    if (f == nb && nb < p.n_frames) p.res->status = p.status[nb];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        rec += __shfl_xor_sync(0xffffffffu, rec, m);
        drop += __shfl_xor_sync(0xffffffffu, drop, m);
    }
    __shared__ unsigned long long s_rec[8], s_drop[8];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_rec[w] = rec;
        s_drop[w] = drop;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long r = 0, d = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            r += s_rec[i];
            d += s_drop[i];
        }
        if (r) atomicAdd(&p.res->received, r);
        if (d) atomicAdd(&p.res->dropped, d);
    }
}

struct TraceOut {
    float* t[SSPSD_MAX_TRACES];
    unsigned long long cap;  // floats available per trace (bounds-checking build)
};

__device__ __forceinline__ float4 adc_word(uint2 w, bool dac)
{
    // i16 as f32 * (4.096 * 2.5 / 32768), data.rs:31-35 (f32 bits 0x39a3d70b); DAC words are offset
    // binary: wrapping_add(i16::MIN) flips bit 15, data.rs:64,75
    const float k = __int_as_float(0x39a3d70b);
    unsigned int flip = dac ? 0x80008000u : 0u;
    unsigned int a = w.x ^ flip, b = w.y ^ flip;
    float4 o;
    o.x = __fmul_rn((float)(short)(a & 0xffffu), k);
    o.y = __fmul_rn((float)(short)(a >> 16), k);
    o.z = __fmul_rn((float)(short)(b & 0xffffu), k);
    o.w = __fmul_rn((float)(short)(b >> 16), k);
    return o;
}

// AdcDac fast path: frames base 8-byte aligned, stride a multiple of 8.  Flat scan of the frame buffer:
// one 8-byte word per lane and step (a warp reads 256 contiguous bytes).  A word that falls into a payload
// is four i16 samples of one channel and becomes one float4 store; neighbouring lanes hold the two halves
// of one 8-sample group, so every store instruction writes whole 32-byte sectors and a warp writes 64-byte
// runs per channel.  A CTA owns a contiguous 32 KiB byte range: the frame index / offset within the frame
// is divided out once per thread and then advanced incrementally, and the loads of 8 steps are issued
// before the first store so that 8 requests per thread are in flight (the first version was stalled on
// long_scoreboard with one load in flight and wrote half sectors, profiles/r01_ncu_summary.md).
constexpr int ADC_NT = 256;
constexpr int ADC_ITERS = 16;  // 8-byte words per thread
constexpr int ADC_MLP = 8;     // loads in flight per thread
__global__ void __launch_bounds__(ADC_NT) adcdac_flat_kernel(const uint8_t* __restrict__ frames,
                                                              unsigned long long n_words8, unsigned long long stride,
                                                              unsigned long long frame_len,
                                                              const DecodeResult* __restrict__ res, TraceOut out)
{
    const unsigned long long nb = res->first_bad;
    const unsigned int spf = res->batches * 8u;  // samples per trace per frame
    const uint2* __restrict__ words = reinterpret_cast<const uint2*>(frames);
    const unsigned long long j0 = (unsigned long long)blockIdx.x * (ADC_NT * ADC_ITERS) + threadIdx.x;
    unsigned long long f = (j0 * 8ull) / stride;
    unsigned long long r = j0 * 8ull - f * stride;
    const unsigned int step = ADC_NT * 8u;  // bytes between two words of one thread
#pragma unroll 1
    for (int it0 = 0; it0 < ADC_ITERS; it0 += ADC_MLP) {
        uint2 v[ADC_MLP];
#pragma unroll
        for (int u = 0; u < ADC_MLP; ++u) {
            const unsigned long long j = j0 + (unsigned long long)(it0 + u) * ADC_NT;
            v[u] = j < n_words8 ? __ldg(words + j) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < ADC_MLP; ++u) {
            const unsigned long long j = j0 + (unsigned long long)(it0 + u) * ADC_NT;
            if (j < n_words8 && f < nb && r >= SSPSD_HEADER_SIZE && r < frame_len) {
                const unsigned int w = (unsigned int)(r - SSPSD_HEADER_SIZE) >> 3;  // 8-byte word of the payload
                const unsigned int b = w >> 3, c = (w >> 1) & 3u, h = w & 1u;
                SSPSD_ASSERT(f * spf + b * 8u + h * 4u + 4u <= out.cap);
                *reinterpret_cast<float4*>(out.t[c] + f * spf + b * 8u + h * 4u) = adc_word(v[u], c >= 2);
            }
            r += step;
            while (r >= stride) {
                r -= stride;
                ++f;
            }
        }
    }
}

// generic path: one thread per batch, byte loads (any alignment, all formats)
__global__ void decode_generic_kernel(const uint8_t* __restrict__ frames, unsigned long long stride,
                                      const DecodeResult* __restrict__ res, TraceOut out)
{
    const unsigned long long nb = res->first_bad;
    const unsigned int fmt = res->format, bat = res->batches;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bat == 0 || i >= nb * bat) return;
    unsigned long long f = i / bat;
    unsigned int b = (unsigned int)(i - f * bat);
    const uint8_t* d = frames + f * stride + SSPSD_HEADER_SIZE + (unsigned long long)b * batch_bytes(fmt);
    const unsigned long long o = f * bat + b;
    SSPSD_ASSERT((fmt == SSPSD_FORMAT_ADCDAC ? o * 8 + 8 : o + 1) <= out.cap);
    if (fmt == SSPSD_FORMAT_ADCDAC) {
        const float k = __int_as_float(0x39a3d70b);
        for (int c = 0; c < 4; ++c)
            for (int s = 0; s < 8; ++s) {
                unsigned int u = (unsigned int)d[16 * c + 2 * s] | ((unsigned int)d[16 * c + 2 * s + 1] << 8);
                if (c >= 2) u ^= 0x8000u;
                out.t[c][o * 8 + s] = __fmul_rn((float)(short)u, k);
            }
    } else if (fmt == SSPSD_FORMAT_FLS) {
        // data.rs:97-139; batch = [[i32;7];2]
        const float inv31 = __int_as_float(0x30000000);  // 1 / (i32::MAX as f32) = 2^-31
        const float kap = __int_as_float(0x38c90fdb);    // TAU / 65536
        float re = __int2float_rn((int)rd_u32(d)), im = __int2float_rn((int)rd_u32(d + 4));
        out.t[0][o] = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))), inv31);
        long long ph = (long long)((unsigned long long)rd_u32(d + 8) | ((unsigned long long)rd_u32(d + 12) << 32));
        out.t[1][o] = __fmul_rn(__ll2float_rn(ph), kap);
        out.t[2][o] = __fmul_rn(__int2float_rn((int)rd_u32(d + 28)), inv31);  // x / 2^31 is exact either way
        out.t[3][o] = __fmul_rn(__int2float_rn((int)rd_u32(d + 32)), inv31);
    } else if (fmt == SSPSD_FORMAT_THERMOSTAT_EEM) {
        // data.rs:154-163: f32 words 0, 8, 13, 16 of 20
        out.t[0][o] = __uint_as_float(rd_u32(d));
        out.t[1][o] = __uint_as_float(rd_u32(d + 32));
        out.t[2][o] = __uint_as_float(rd_u32(d + 52));
        out.t[3][o] = __uint_as_float(rd_u32(d + 64));
    } else {
        // Mpll, data.rs:178-211
        const float kph = __int_as_float(0x30c90fdb);  // TAU / 2^32
        const float kfr = __int_as_float(0x34435000);  // 1 / 1.28e-3 / 2^32
        const float kam = __int_as_float(0x3083126e);  // 10.24 / 10 * 2 * 2 / 2^32
        out.t[0][o] = __fmul_rn(__int2float_rn((int)rd_u32(d + 16)), kph);
        out.t[1][o] = __fmul_rn(__int2float_rn((int)rd_u32(d + 20)), kfr);
        float re = __int2float_rn((int)rd_u32(d)), im = __int2float_rn((int)rd_u32(d + 4));
        out.t[2][o] = __fmul_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im))), kam);
    }
}

}  // namespace sspsd
