// Microbenchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2 / add.rn.f32x2, sm_100+) issue and
// pipe throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

constexpr int ITERS = 4096, ACC = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s)
{
    float2 a[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 b = make_float2(s, s * 0.5f), c = make_float2(1e-7f, 2e-7f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) {  // scalar FFMA x2
                a[i].x = fmaf(a[i].x, b.x, c.x);
                a[i].y = fmaf(a[i].y, b.y, c.y);
            } else if (MODE == 1) {  // packed FFMA2
                a[i] = fma2(a[i], b, c);
            } else if (MODE == 2) {  // scalar FADD x2
                a[i].x = __fadd_rn(a[i].x, c.x);
                a[i].y = __fadd_rn(a[i].y, c.y);
            } else if (MODE == 3) {  // packed FADD2
                a[i] = add2(a[i], c);
            } else if (MODE == 4) {  // scalar FFMA x2 with all-register operands varying
                a[i].x = fmaf(a[i].x, a[(i + 1) % ACC].y, c.x);
                a[i].y = fmaf(a[i].y, a[(i + 1) % ACC].x, c.y);
            } else {  // packed, register operands
                a[i] = fma2(a[i], a[(i + 1) % ACC], c);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, float* d)
{
    const int grid = 148 * 8;
    k<MODE><<<grid, 256>>>(d, 0.999f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(d, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    double ops = (double)grid * 256 * ITERS * ACC * 2;  // scalar-equivalent FP32 instructions (lane ops)
    printf("%-28s %8.3f ms  %7.2f T lane-ops/s  (%5.1f per clk per SM @1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main()
{
    float* d;
    cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    run<0>("scalar FFMA (imm/const b)", d);
    run<1>("packed FFMA2", d);
    run<2>("scalar FADD", d);
    run<3>("packed FADD2", d);
    run<4>("scalar FFMA 3-reg", d);
    run<5>("packed FFMA2 3-reg", d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return 0;
}
