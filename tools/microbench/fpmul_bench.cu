// fpmul_bench.cu -- throughput of the Montgomery product variants on the integer pipe (round 2):
//   cios   : saturated 32-bit limbs, carry-chained IMAD.WIDE.U32.X / IMAD.HI (the round-1 product, fp_mul_cios)
//   unsat  : unsaturated radix 2^30 / 2^29, plain IMAD.WIDE.U32 column sums (fp_mul_unsat, zkm_fpmul_u.cuh)
//   sqr    : dedicated squaring on the same columns (fp_sqr_unsat)
// and of the XYZZ mixed addition built on them.  Also checks on the device that both products return the same
// bytes for 2^20 random operand pairs per field.  "GMAD_equiv_s" = products/s x (2 n^2 + n), the canonical
// 32-bit-limb MAD count of SURVEY 8d, so the numbers are comparable with profiles/imad_peak_r1.jsonl.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o fpmul_bench fpmul_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../zkmember_b200/csrc/zkm_curve.cuh"
using namespace zkm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <class F>
__device__ F rand_elem(uint32_t seed) {   // a raw representative below 2^(BITS-1) < p
    F x;
    for (int i = 0; i < F::N; i++) x.l[i] = mix(seed * 977u + i * 0x9e3779b9u);
    constexpr int topbits = F::Params::BITS - 1 - 32 * (F::N - 1);
    x.l[F::N - 1] &= (topbits >= 32) ? 0xffffffffu : ((1u << topbits) - 1u);
    return x;
}

template <class F, int V>
__global__ void __launch_bounds__(128) k_modmul(uint32_t* out, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x = rand_elem<F>(2 * tid + 1), y = rand_elem<F>(2 * tid + 2);
    for (int it = 0; it < iters; it++) {
        if (V == 0) { x = fp_mul_cios(x, y); y = fp_mul_cios(y, x); }
        if (V == 1) { x = fp_mul_unsat(x, y); y = fp_mul_unsat(y, x); }
        if (V == 2) { x = fp_sqr_unsat(x); y = fp_sqr_unsat(y); }
    }
    uint32_t s = 0;
    for (int i = 0; i < F::N; i++) s ^= x.l[i] ^ y.l[i];
    out[tid] = s;
}
// four independent chains per thread (ILP inside one warp, as in the kernels that batch independent products)
template <class F, int V>
__global__ void __launch_bounds__(128) k_modmul4(uint32_t* out, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    F x[4], y = rand_elem<F>(tid + 77);
    for (int j = 0; j < 4; j++) x[j] = rand_elem<F>(4 * tid + j);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 4; j++) x[j] = V == 0 ? fp_mul_cios(x[j], y) : fp_mul_unsat(x[j], y);
    }
    uint32_t s = 0;
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < F::N; i++) s ^= x[j].l[i];
    out[tid] = s;
}
template <class F>
__global__ void __launch_bounds__(128) k_madd(uint32_t* out, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::identity();
    F x = rand_elem<F>(2 * tid + 1), y = rand_elem<F>(2 * tid + 2);
    for (int it = 0; it < iters; it++) { xyzz_madd(acc, x, y); x.l[0] ^= it; }
    uint32_t s = 0;
    for (int i = 0; i < F::N; i++) s ^= acc.X.l[i] ^ acc.Y.l[i] ^ acc.ZZ.l[i] ^ acc.ZZZ.l[i];
    out[tid] = s;
}
template <class F>
__global__ void k_check(uint32_t* bad, uint32_t n) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    F x = rand_elem<F>(3 * tid + 1), y = rand_elem<F>(3 * tid + 2);
    if (tid % 5 == 0) for (int i = 0; i < F::N; i++) x.l[i] = F::Params::mod(i) - (i == 0 ? 1 + tid % 7 : 0);   // p - small
    F a = fp_mul_cios(x, y), b = fp_mul_unsat(x, y), c = fp_mul_cios(x, x), d = fp_sqr_unsat(x);
    if (a != b || c != d) atomicAdd(bad, 1u);
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

static int g_sms;
static uint32_t* g_out;

template <class F>
static void bench_field(const char* name) {
    const double mads = 2.0 * F::N * F::N + F::N;
    uint32_t* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    k_check<F><<<(1 << 20) / 256, 256>>>(bad, 1u << 20);
    uint32_t hbad = 0; CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"%s\", \"pairs\": %d, \"mismatches\": %u}\n", name, 1 << 20, hbad);
    const int iters = F::N >= 24 ? 128 : 512;
    for (int warps_per_sm : {4, 8, 12, 16, 24, 32}) {
        int threads = 128, blocks = g_sms * warps_per_sm * 32 / threads;
        double n_mul = (double)blocks * threads * iters * 2;
        float t0 = time_kernel([&] { k_modmul<F, 0><<<blocks, threads>>>(g_out, iters); });
        float t1 = time_kernel([&] { k_modmul<F, 1><<<blocks, threads>>>(g_out, iters); });
        float t2 = time_kernel([&] { k_modmul<F, 2><<<blocks, threads>>>(g_out, iters); });
        printf("{\"bench\": \"modmul\", \"field\": \"%s\", \"warps_per_sm\": %d, \"cios_Gmul_s\": %.2f, \"unsat_Gmul_s\": %.2f, "
               "\"sqr_Gmul_s\": %.2f, \"cios_GMAD_equiv_s\": %.1f, \"unsat_GMAD_equiv_s\": %.1f, \"speedup\": %.3f}\n",
               name, warps_per_sm, n_mul / t0 / 1e6, n_mul / t1 / 1e6, n_mul / t2 / 1e6, n_mul * mads / t0 / 1e6,
               n_mul * mads / t1 / 1e6, t0 / t1);
        if (warps_per_sm <= 16 && F::N <= 12) {
            double n4 = (double)blocks * threads * iters * 4;
            float u0 = time_kernel([&] { k_modmul4<F, 0><<<blocks, threads>>>(g_out, iters); });
            float u1 = time_kernel([&] { k_modmul4<F, 1><<<blocks, threads>>>(g_out, iters); });
            printf("{\"bench\": \"modmul_ilp4\", \"field\": \"%s\", \"warps_per_sm\": %d, \"cios_Gmul_s\": %.2f, \"unsat_Gmul_s\": %.2f, "
                   "\"unsat_GMAD_equiv_s\": %.1f, \"speedup\": %.3f}\n", name, warps_per_sm, n4 / u0 / 1e6, n4 / u1 / 1e6,
                   n4 * mads / u1 / 1e6, u0 / u1);
            double n_madd = (double)blocks * threads * iters;
            float tm = time_kernel([&] { k_madd<F><<<blocks, threads>>>(g_out, iters); });
            printf("{\"bench\": \"xyzz_madd\", \"field\": \"%s\", \"warps_per_sm\": %d, \"Gmadd_s\": %.3f, \"GMAD_equiv_s\": %.1f}\n",
                   name, warps_per_sm, n_madd / tm / 1e6, n_madd * 10 * mads / tm / 1e6);
        }
    }
    CK(cudaFree(bad));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    CK(cudaMalloc(&g_out, sizeof(uint32_t) * g_sms * 64 * 1024));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, g_sms, clk_khz);
    bench_field<Bls12_381_Fq>("bls12_381_fq");
    bench_field<Bls12_381_Fr>("bls12_381_fr");
    bench_field<Bn254_Fq>("bn254_fq");
    bench_field<Bw6_761_Fr>("bw6_761_fr");
    bench_field<Bw6_761_Fq>("bw6_761_fq");
    return 0;
}
