// int_pipe_peak.cu -- issue rates of the integer instructions a Montgomery product is made of, measured so
// that ptxas CANNOT rewrite them (round 1's imad_peak.cu multiplied two loop-invariant kernel parameters:
// ptxas hoisted the product and the "imad_wide" loop became IADD3/IADD3.X pairs -- its 62 /clk/SM was the
// rate of additions, not of IMAD.WIDE).  Here every multiplicand depends on the accumulator it feeds.
// `cuobjdump -sass` of this file is committed next to the numbers (profiles/int_pipe_peak_r2.sass.txt).
//
// Output: JSON lines, ops per clock per SM at the maximum SM clock, for 8/16/32 warps per SM.
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o int_pipe_peak int_pipe_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITER = 2048;
constexpr int ILP = 8;

// MODE 0: mad.wide.u32 d, a, b, c in PTX: ptxas SPLITS it into IMAD.WIDE.U32 d, a, b, RZ + IADD3 / IADD3.X
// MODE 14: the same written in C++ (mul.wide.u32 + add.s64): ptxas FUSES it into IMAD.WIDE.U32 d, a, b, c
// MODE 1: IMAD.WIDE.U32 d, a, b, RZ  (mul.wide)                          acc = lo(acc) * hi(acc)
// MODE 2: IMAD (mad.lo)                                                   x = x * b + c
// MODE 3: IMAD.HI                                                         x = hi(x * b) + c
// MODE 4: carry chain mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X          (the saturated CIOS inner step)
// MODE 5: IADD3 (add)            MODE 6: SHF (funnel shift)              MODE 7: LOP3 (and/xor)
// MODE 8: DFMA                   MODE 9: one IMAD.WIDE + one DFMA per step (are the pipes independent?)
// MODE 10/11/12: one IMAD.WIDE + 1/2/3 IADD3 per step (how many ALU operations hide behind a wide MAD)
// MODE 13: one IMAD.WIDE + one IMAD (lo) per step
template <int MODE>
__global__ void __launch_bounds__(256) k_rate(uint32_t* out, uint32_t b, uint32_t c) {
    uint64_t acc[ILP];
    uint32_t x[ILP], y[ILP];
    double d[ILP];
    for (int i = 0; i < ILP; i++) {
        acc[i] = ((uint64_t)(threadIdx.x + i) << 32) | (threadIdx.x * 7 + i + 1);
        x[i] = threadIdx.x + i + 1;
        y[i] = threadIdx.x * 3 + i;
        d[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    }
    const double db = 1.0 + 1e-12 * b, dc = 1e-15 * c;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0 || MODE >= 9) {
                uint32_t lo = (uint32_t)acc[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo), "r"(b));
            }
            if (MODE == 14) acc[i] += (uint64_t)(uint32_t)acc[i] * b;     // C++ form: ptxas fuses mul.wide + add.s64
            if (MODE == 1) {
                uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32) | 1u;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(lo), "r"(hi));
            }
            if (MODE == 2 || MODE == 13) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
            if (MODE == 3) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
            if (MODE == 5 || MODE == 10 || MODE == 11 || MODE == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (MODE == 11 || MODE == 12) asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(x[i]));
            if (MODE == 12) asm volatile("sub.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
            if (MODE == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y[i]));
            if (MODE == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(c));
            if (MODE == 8 || MODE == 9) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(db), "d"(dc));
        }
        if (MODE == 4) {   // one chain of 2 ILP limbs: multiplicand = previous low word of the chain
            uint32_t* w = reinterpret_cast<uint32_t*>(acc);
            uint32_t m = w[0] | 1u;
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(w[0]) : "r"(m), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(w[1]) : "r"(m), "r"(b));
#pragma unroll
            for (int i = 1; i < ILP; i++) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(w[2 * i]) : "r"(m), "r"(b));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(w[2 * i + 1]) : "r"(m), "r"(b));
            }
        }
    }
    uint64_t s = 0;
    double ds = 0;
    for (int i = 0; i < ILP; i++) { s ^= acc[i] ^ x[i] ^ ((uint64_t)y[i] << 7); ds += d[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32)) ^ (uint32_t)__double_as_longlong(ds);
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

static int g_sms, g_clk;
static uint32_t* g_out;

template <int MODE>
static void run(const char* name, const char* sass, double ops_per_step) {
    for (int warps : {8, 16, 32}) {
        int threads = 256, blocks = g_sms * warps * 32 / threads;
        float t = time_kernel([&] { k_rate<MODE><<<blocks, threads>>>(g_out, 0x9e3779b1u, 12345u); });
        double steps = (double)blocks * threads * ITER * ILP;
        printf("{\"bench\": \"%s\", \"sass\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"steps_per_clk_per_sm\": %.2f, "
               "\"Gops_s\": %.1f, \"ops_per_step\": %.0f}\n", name, sass, warps, t, steps / (t * 1e-3) / g_sms / (g_clk * 1e3),
               steps * ops_per_step / t / 1e6, ops_per_step);
    }
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    CK(cudaDeviceGetAttribute(&g_clk, cudaDevAttrClockRate, 0));
    CK(cudaMalloc(&g_out, sizeof(uint32_t) * g_sms * 64 * 1024));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, g_sms, g_clk);
    run<0>("imad_wide_acc64", "IMAD.WIDE.U32 Rd, Ra, Rb, Rc", 1);
    run<14>("imad_wide_acc64_fused", "IMAD.WIDE.U32 Rd, Ra, Rb, Rc (64-bit addend)", 1);
    run<1>("mul_wide", "IMAD.WIDE.U32 Rd, Ra, Rb, RZ", 1);
    run<2>("imad_lo", "IMAD", 1);
    run<3>("imad_hi", "IMAD.HI.U32", 1);
    run<4>("imad_wide_carry_chain", "IMAD.WIDE.U32.X (mad.lo.cc + madc.hi.cc)", 1);
    run<5>("iadd", "IADD3", 1);
    run<6>("shf", "SHF.R.W", 1);
    run<7>("lop3", "LOP3.LUT", 1);
    run<8>("dfma", "DFMA", 1);
    run<9>("imad_wide+dfma", "IMAD.WIDE.U32 + DFMA", 2);
    run<10>("imad_wide+1alu", "IMAD.WIDE.U32 + 1 IADD3", 2);
    run<11>("imad_wide+2alu", "IMAD.WIDE.U32 + 2 IADD3", 3);
    run<12>("imad_wide+3alu", "IMAD.WIDE.U32 + 3 IADD3", 4);
    run<13>("imad_wide+imad_lo", "IMAD.WIDE.U32 + IMAD", 2);
    return 0;
}
