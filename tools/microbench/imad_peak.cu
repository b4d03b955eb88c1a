// imad_peak.cu -- integer-pipe microbenchmark for the MSM roofline denominator
// ("Peak MAD rate = measured by a mad.wide.u32 microbenchmark on the box").
// Measures per-SM issue rates of IMAD / IMAD.HI / IMAD.WIDE / carry-chained
// IMAD.WIDE.X / IADD3, and the modmul rate of the library's own fp_mul.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o imad_peak imad_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../zkmember_b200/csrc/zkm_curve.cuh"
using namespace zkm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITER = 4096;
constexpr int ILP = 8;

__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_hi(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_wide(uint32_t* out, uint32_t a, uint32_t b) {
    unsigned long long acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
    }
    unsigned long long s = 0;
    for (int i = 0; i < ILP; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// carry-chained lo/hi pairs (what fp_mul issues): ILP independent... a chain is serial by the
// carry, so ILP comes from warps only -- this is the realistic shape.
__global__ void k_chain(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[2 * ILP];
    for (int i = 0; i < 2 * ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
        acc[0] = ptx::mad_lo_cc(a, b, acc[0]);
        acc[1] = ptx::madc_hi_cc(a, b, acc[1]);
#pragma unroll
        for (int i = 1; i < ILP; i++) {
            acc[2 * i] = ptx::madc_lo_cc(a, b, acc[2 * i]);
            acc[2 * i + 1] = ptx::madc_hi_cc(a, b, acc[2 * i + 1]);
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 2 * ILP; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(a));
    }
    uint32_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
template <class F>
__global__ void k_modmul(uint32_t* out, int iters) {
    F x, y;
    for (int i = 0; i < F::N; i++) { x.l[i] = F::Params::one(i) + threadIdx.x; y.l[i] = F::Params::r2(i); }
    x.l[F::N - 1] &= 0x0fffffff; 
    for (int it = 0; it < iters; it++) { x = x * y; y = y * x; }
    uint32_t s = 0;
    for (int i = 0; i < F::N; i++) s ^= x.l[i] ^ y.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
__global__ void k_madd(uint32_t* out, int iters) {
    XYZZ<F> acc = XYZZ<F>::identity();
    F x, y;
    for (int i = 0; i < F::N; i++) { x.l[i] = F::Params::one(i) + threadIdx.x; y.l[i] = F::Params::r2(i); }
    x.l[F::N - 1] &= 0x0fffffff;
    for (int it = 0; it < iters; it++) { xyzz_madd(acc, x, y); x.l[0] ^= it; }
    uint32_t s = 0;
    for (int i = 0; i < F::N; i++) s ^= acc.X.l[i] ^ acc.Y.l[i] ^ acc.ZZ.l[i] ^ acc.ZZZ.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static float time_kernel(K launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    uint32_t* out; CK(cudaMalloc(&out, sizeof(uint32_t) * sms * 64 * 1024));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, clk_khz);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
        if (blocks < 1) blocks = 1;
        double n_inst = (double)blocks * threads * ITER * ILP;
        float t;
        t = time_kernel([&] { k_imad_lo<<<blocks, threads>>>(out, 3, 5); });
        printf("{\"bench\": \"imad_lo\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", warps_per_sm, t, n_inst / t / 1e6, n_inst / (t * 1e-3) / sms / (clk_khz * 1e3));
        t = time_kernel([&] { k_imad_hi<<<blocks, threads>>>(out, 3, 5); });
        printf("{\"bench\": \"imad_hi\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", warps_per_sm, t, n_inst / t / 1e6, n_inst / (t * 1e-3) / sms / (clk_khz * 1e3));
        t = time_kernel([&] { k_imad_wide<<<blocks, threads>>>(out, 3, 5); });
        printf("{\"bench\": \"imad_wide\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", warps_per_sm, t, n_inst / t / 1e6, n_inst / (t * 1e-3) / sms / (clk_khz * 1e3));
        t = time_kernel([&] { k_chain<<<blocks, threads>>>(out, 3, 5); });
        printf("{\"bench\": \"chain_wide_x\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops_wide\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", warps_per_sm, t, n_inst / t / 1e6, n_inst / (t * 1e-3) / sms / (clk_khz * 1e3));
        t = time_kernel([&] { k_iadd3<<<blocks, threads>>>(out, 3, 5); });
        printf("{\"bench\": \"iadd\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.1f, \"per_sm_per_clk_at_max\": %.2f}\n", warps_per_sm, t, n_inst / t / 1e6, n_inst / (t * 1e-3) / sms / (clk_khz * 1e3));
    }
    for (int warps_per_sm : {4, 8, 12, 16, 32}) {
        int threads = 128, blocks = sms * warps_per_sm * 32 / threads;
        int iters = 512;
        double n_mul = (double)blocks * threads * iters * 2;
        float t;
        t = time_kernel([&] { k_modmul<Bls12_381_Fr><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"modmul_fr256\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gmodmul_s\": %.2f, \"GMAD_s\": %.1f}\n", warps_per_sm, t, n_mul / t / 1e6, n_mul * 136 / t / 1e6);
        t = time_kernel([&] { k_modmul<Bls12_381_Fq><<<blocks, threads>>>(out, iters); });
        printf("{\"bench\": \"modmul_fq384\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gmodmul_s\": %.2f, \"GMAD_s\": %.1f}\n", warps_per_sm, t, n_mul / t / 1e6, n_mul * 300 / t / 1e6);
        if (warps_per_sm <= 16) {
            double n_madd = (double)blocks * threads * iters;
            t = time_kernel([&] { k_madd<Bls12_381_Fq><<<blocks, threads>>>(out, iters); });
            printf("{\"bench\": \"xyzz_madd_fq384\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gmadd_s\": %.3f, \"GMAD_s\": %.1f}\n", warps_per_sm, t, n_madd / t / 1e6, n_madd * 3000 / t / 1e6);
            t = time_kernel([&] { k_madd<Bn254_Fq><<<blocks, threads>>>(out, iters); });
            printf("{\"bench\": \"xyzz_madd_fq256\", \"warps_per_sm\": %d, \"ms\": %.4f, \"Gmadd_s\": %.3f, \"GMAD_s\": %.1f}\n", warps_per_sm, t, n_madd / t / 1e6, n_madd * 1360 / t / 1e6);
        }
    }
    return 0;
}
