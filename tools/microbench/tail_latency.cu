// tail_latency.cu -- latency of the dependent group operations the serial tails of an MSM are made of
// (window reduction, fold levels, Horner combine): cycles per operation of a chain of N dependent operations
// on a lone warp (and with 4 / 8 warps on the SM), BLS12-381 G1.
//
//   fp_mul            one dependent Montgomery product (12 limbs)
//   add / add_ni      xyzz_add inlined / out of line, one lane per chain
//   add_quad          xyzz_add_quad (four lanes per chain, zkm_msm_quad.cuh)
//   dbl / dbl_quad    xyzz_dbl inlined / xyzz_dbl_quad_inl
//   madd              xyzz_madd (affine operand)
//
// Output: JSON lines {op, warps, cycles_per_op, us_per_op}.  Every chain ends in a store, so nothing is elided.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I../../zkmember_b200/csrc -I../../include \
//        -o tail_latency tail_latency.cu
#include <cstdio>
#include <cstdlib>

#include "zkm_msm_curve.cuh"

using namespace zkm;
typedef Bls12_381_Fq F;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { OP_MUL = 0, OP_ADD, OP_ADD_NI, OP_ADD_QUAD, OP_DBL, OP_DBL_QUAD, OP_MADD, OP_DBL_NI, OP_DBL_QUAD_NI, N_OPS };
static const char* OP_NAME[N_OPS] = {"fp_mul", "xyzz_add", "xyzz_add_ni", "xyzz_add_quad", "xyzz_dbl", "xyzz_dbl_quad_inl",
                                     "xyzz_madd", "xyzz_dbl_ni", "xyzz_dbl_quad"};

template <int OP>
__global__ void __launch_bounds__(256) k_chain(int iters, long long* cycles, uint32_t* sink) {
    F gx, gy;
    load_generator<G1Bls>(gx, gy);
    XYZZ<F> q = xyzz_from_affine(gx, gy);
    XYZZ<F> p;
    xyzz_mdbl(p, gx, gy);
    xyzz_add_ni(p, q);          // p = 3 G: generic Z
    XYZZ<F> q2 = p;
    xyzz_dbl_ni(q2);            // q2 = 6 G, projective operand for the additions
    const int lane4 = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    F acc = gx;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (OP == OP_MUL) acc = acc * gy;
        if (OP == OP_ADD) xyzz_add(p, q2);
        if (OP == OP_ADD_NI) xyzz_add_ni(p, q2);
        if (OP == OP_ADD_QUAD) xyzz_add_quad(p, q2, lane4, mask);
        if (OP == OP_DBL) xyzz_dbl(p);
        if (OP == OP_DBL_NI) xyzz_dbl_ni(p);
        if (OP == OP_DBL_QUAD) xyzz_dbl_quad_inl(p, lane4, mask);
        if (OP == OP_DBL_QUAD_NI) xyzz_dbl_quad(p, lane4, mask);
        if (OP == OP_MADD) xyzz_madd(p, gx, gy);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t h = 0;
    for (int i = 0; i < 12; i++) h ^= acc.l[i] ^ p.X.l[i] ^ p.Y.l[i] ^ p.ZZ.l[i] ^ p.ZZZ.l[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = h;
}

template <int OP>
static void run(int warps, int iters, double mhz) {
    long long* d_cyc;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, 8));
    CK(cudaMalloc(&d_sink, 4 * 32 * warps));
    for (int rep = 0; rep < 2; rep++) k_chain<OP><<<1, 32 * warps>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"op\": \"%s\", \"warps\": %d, \"iters\": %d, \"cycles_per_op\": %.1f, \"us_per_op\": %.3f}\n", OP_NAME[OP], warps,
           iters, (double)c / iters, (double)c / iters / mhz);
    CK(cudaFree(d_cyc));
    CK(cudaFree(d_sink));
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, dev));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const double mhz = khz / 1000.0;
    printf("{\"device\": \"%s\", \"sm_mhz\": %.0f}\n", pr.name, mhz);
    const int warps_list[3] = {1, 4, 8};
    for (int wi = 0; wi < 3; wi++) {
        const int w = warps_list[wi];
        run<OP_MUL>(w, 2000, mhz);
        run<OP_ADD>(w, 300, mhz);
        run<OP_ADD_NI>(w, 300, mhz);
        run<OP_ADD_QUAD>(w, 300, mhz);
        run<OP_DBL>(w, 300, mhz);
        run<OP_DBL_NI>(w, 300, mhz);
        run<OP_DBL_QUAD>(w, 300, mhz);
        run<OP_DBL_QUAD_NI>(w, 300, mhz);
        run<OP_MADD>(w, 300, mhz);
    }
    return 0;
}
