// Issue rate of the packed 16-bit integer instructions the max-log-MAP kernel is made of (development aid).
// One block per SM, 1024 threads (8 warps per scheduler), ILP 8 independent chains per thread, N iterations.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return __vadd2(a, b);                       // VIADD.16x2
    if (OP == 1) return __viaddmax_s16x2(a, b, c);           // VIADDMNMX.S16x2
    if (OP == 2) return __vmaxs2(a, b);                      // VIMNMX.S16x2
    if (OP == 3) return a + b;                               // IADD3
    if (OP == 4) return __byte_perm(a, b, 0x5410);           // PRMT
    if (OP == 5) return (a & b) ^ c;                         // LOP3
    if (OP == 6) return a * b + c;                           // IMAD
    if (OP == 7) return max((int)a, (int)b);                 // VIMNMX (32-bit)
    return __viaddmax_s32(a, b, c);                          // VIADDMNMX (32-bit)
}
template <int OP> __global__ void k(uint32_t *out, int n, uint32_t s0, uint32_t s1)
{
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = s0 + threadIdx.x * 8 + i;
    const uint32_t b = s1 | 1u, c = s0 ^ 0x1234u;
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = op<OP>(x[i], b + i, c);
    }
    uint32_t r = 0;
    for (int i = 0; i < 8; i++) r ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> void run(const char *name, uint32_t *d, int sms)
{
    const int n = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<sms, 1024>>>(d, 16, 1, 2);
    cudaEventRecord(a);
    k<OP><<<sms, 1024>>>(d, n, 1, 2);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_inst = (double)n * 64 * 32;            // per SM: 32 warps x 64 instructions per iteration
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-22s %.3f warp-instructions / cycle / SM  (%.2f per scheduler)\n", name, warp_inst / cycles, warp_inst / cycles / 4);
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *d; cudaMalloc(&d, sizeof(uint32_t) * sms * 1024);
    run<0>("VIADD.16x2", d, sms); run<1>("VIADDMNMX.S16x2", d, sms); run<2>("VIMNMX.S16x2", d, sms); run<3>("IADD3", d, sms);
    run<4>("PRMT", d, sms); run<5>("LOP3", d, sms); run<6>("IMAD", d, sms); run<7>("VIMNMX (s32)", d, sms); run<8>("VIADDMNMX (s32)", d, sms);
    return 0;
}
