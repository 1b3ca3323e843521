// Microbenchmark (B200): issue/pipe throughput of the instruction classes the GN pixel loop is made of.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 2048;
constexpr int NACC = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, uint32_t iseed) {
    float a[NACC]; float2 a2[NACC]; uint32_t u[NACC];
    const float x = seed + threadIdx.x * 1e-6f, y = 1.0f - 1e-7f * threadIdx.x;
    for (int i = 0; i < NACC; ++i) { a[i] = x + i; a2[i] = make_float2(x + i, y + i); u[i] = iseed + i + threadIdx.x; }
    const float2 x2 = make_float2(x, y), y2 = make_float2(y, x);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], x, y);                                   // FFMA
            if (MODE == 1) a2[i] = __ffma2_rn(a2[i], x2, y2);                         // FFMA2
            if (MODE == 2) { a[i] = fmaf(a[i], x, y); u[i] = (u[i] & iseed) ^ (u[i] >> 3); }            // FFMA + LOP3-ish (SHF+LOP3)
            if (MODE == 3) { a2[i] = __ffma2_rn(a2[i], x2, y2); u[i] = (u[i] & iseed) ^ (u[i] >> 3); }
            if (MODE == 4) u[i] = (u[i] & iseed) ^ (u[i] >> 3);                        // ALU only (SHF + LOP3)
            if (MODE == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));  // MUFU
            if (MODE == 6) a[i] = (float)__float2int_rd(a[i]) + x;                    // F2I + I2FP + FADD
            if (MODE == 7) a[i] = a[i] + x;                                           // FADD
            if (MODE == 8) a2[i] = __fadd2_rn(a2[i], x2);                             // FADD2
            if (MODE == 9) a[i] = a[i] * x;                                           // FMUL
            if (MODE == 10) u[i] = __byte_perm(u[i], iseed, 0x7643) + 1;              // PRMT + IADD
            if (MODE == 11) { a[i] = fmaf(a[i], x, y); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a2[i].x)); }   // FFMA + MUFU
            if (MODE == 12) { a2[i] = __ffma2_rn(a2[i], x2, y2); a[i] = fmaf(a[i], x, y); }   // FFMA2 + FFMA
            if (MODE == 13) a[i] = (u[i] > iseed) ? a[i] : x;                          // ISETP+FSEL
        }
    }
    float s = 0; uint32_t su = 0;
    for (int i = 0; i < NACC; ++i) { s += a[i] + a2[i].x + a2[i].y; su += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)su;
}

template <int MODE>
void run(const char* name, int ops_per_inner, float* d) {
    const int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 1.0f, 0x0ff0f0f0u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 1.0f, 0x0ff0f0f0u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inner = (double)blocks * 8 * ITERS * NACC;   // inner-body executions per warp summed
    const double clk = 1.965e9;                                     // nominal; ratios are what matters
    const double per_smsp_cycles = ms * 1e-3 * clk;                 // cycles elapsed
    const double inner_per_smsp = warp_inner / (148.0 * 4.0);
    printf("%-28s %8.3f ms   cycles per inner body per SMSP: %6.3f   (%d listed ops)\n", name, ms, per_smsp_cycles / inner_per_smsp, ops_per_inner);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA", 1, d);
    run<1>("FFMA2", 1, d);
    run<2>("FFMA + SHF + LOP3", 3, d);
    run<3>("FFMA2 + SHF + LOP3", 3, d);
    run<4>("SHF + LOP3", 2, d);
    run<5>("MUFU.RCP", 1, d);
    run<6>("F2I.FLOOR + I2FP + FADD", 3, d);
    run<7>("FADD", 1, d);
    run<8>("FADD2", 1, d);
    run<9>("FMUL", 1, d);
    run<10>("PRMT + IADD", 2, d);
    run<11>("FFMA + MUFU", 2, d);
    run<12>("FFMA2 + FFMA", 2, d);
    run<13>("ISETP + FSEL", 2, d);
    return 0;
}
