// sums_check.cu -- unit check of k_sums_stage1 / k_sums_stage2 (plain sums of the hierarchical reduction's acc arrays)
// against a serial single-thread sum, BLS12-381 G1.  Prints MATCH / MISMATCH per (level, window).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "zkm_msm_curve.cuh"
using namespace zkm;
typedef Bls12_381_Fq F;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_fill(XYZZ<F>* out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F gx, gy;
    load_generator<G1Bls>(gx, gy);
    XYZZ<F> p = xyzz_mul_u64_ni(gx, gy, (uint64_t)(i % 97 == 5 ? 0 : 3 * i + 1));   // some identities
    st_xyzz(out + i, p);
}
__global__ void k_ref(const XYZZ<F>* base, SumJobs jb, uint32_t W, uint64_t* out) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (uint32_t)jb.n * W) return;
    uint32_t L = id / W, w = id % W;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t t = 0; t < jb.per[L]; t++) {
        XYZZ<F> v = ld_xyzz(base + jb.off[L] + (size_t)w * jb.per[L] + t);
        xyzz_add_ni(acc, v);
    }
    write_result<F>(out + (size_t)id * 14, acc);   // 13 words used, 16-byte aligned stride
}
__global__ void k_norm(const XYZZ<F>* sums, uint32_t n, uint64_t* out) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    XYZZ<F> v = ld_xyzz(sums + id);
    write_result<F>(out + (size_t)id * 14, v);
}
// variants of stage 1 for bisection
template <int VAR>
__global__ void __launch_bounds__(128) k_var(const XYZZ<F>* __restrict__ base, SumJobs jb, uint32_t W, XYZZ<F>* __restrict__ slice_sums) {
    __shared__ XYZZ<F> sh[32];
    int L = 0;
    while (L + 1 < jb.n && blockIdx.x >= jb.cta0[L + 1]) L++;
    const uint32_t r = blockIdx.x - jb.cta0[L];
    const uint32_t w = r / jb.nsl[L], sl = r % jb.nsl[L];
    const uint32_t per = jb.per[L];
    const XYZZ<F>* in = base + jb.off[L] + (size_t)w * per;
    const uint32_t lo = sl * ZKM_SUM_SLICE, hi = lo + ZKM_SUM_SLICE < per ? lo + ZKM_SUM_SLICE : per;
    const int q = threadIdx.x & 3;
    const uint32_t quad = threadIdx.x >> 2;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t t = lo + quad; t < hi; t += 32) {
        XYZZ<F> v = ld_xyzz(in + t);
        xyzz_add_quad(acc, v, q, mask);
    }
    XYZZ<F>* dst = slice_sums + jb.sl0[L] + (size_t)w * jb.nsl[L] + sl;
    if (VAR == 0) {            // no tree: quad 0's own sum (right only when the slice has <= 1 element per quad 0)
        if (threadIdx.x == 0) st_xyzz(dst, acc);
    } else if (VAR == 1) {     // tree written out on the shared array itself
        if (q == 0) sh[quad] = acc;
        __syncthreads();
        for (uint32_t s2 = 16; s2 > 0; s2 >>= 1) {
            if (quad < s2) {
                XYZZ<F> a = sh[quad], b = sh[quad + s2];
                xyzz_add_quad(a, b, q, mask);
                if (q == 0) sh[quad] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) st_xyzz(dst, sh[0]);
    } else {                   // only the store to shared and back
        if (q == 0) sh[quad] = acc;
        __syncthreads();
        if (threadIdx.x == 0) st_xyzz(dst, sh[0]);
    }
}
template <int VAR>
static void run_var(const XYZZ<F>* d_in, SumJobs jb, uint32_t W, uint32_t n_cta, XYZZ<F>* d_sl) {
    k_var<VAR><<<n_cta, 128>>>(d_in, jb, W, d_sl);
    uint64_t* d_o;
    CK(cudaMalloc(&d_o, 8 * 14 * W));
    k_norm<<<1, 32>>>(d_sl + jb.sl0[3], W, d_o);
    std::vector<uint64_t> o(14 * W);
    CK(cudaMemcpy(o.data(), d_o, 8 * o.size(), cudaMemcpyDeviceToHost));
    printf("variant %d last level: x0 = %016llx %016llx %016llx\n", VAR, (unsigned long long)o[0], (unsigned long long)o[14], (unsigned long long)o[28]);
}

int main() {
    const uint32_t W = 3;
    const uint32_t pers[4] = {2048, 700, 5, 1};
    SumJobs jb;
    jb.n = 0;
    uint32_t n_cta = 0, n_sl = 0, total = 0;
    for (int L = 0; L < 4; L++) {
        jb.off[L] = total;
        jb.per[L] = pers[L];
        jb.nsl[L] = (pers[L] + ZKM_SUM_SLICE - 1) / ZKM_SUM_SLICE;
        jb.cta0[L] = n_cta;
        jb.sl0[L] = n_sl;
        n_cta += W * jb.nsl[L];
        n_sl += W * jb.nsl[L];
        jb.cta0[L + 1] = n_cta;
        total += W * pers[L];
        jb.n++;
    }
    XYZZ<F>*d_in, *d_sl, *d_sums;
    uint64_t *d_ref, *d_got;
    CK(cudaMalloc(&d_in, sizeof(XYZZ<F>) * total));
    CK(cudaMalloc(&d_sl, sizeof(XYZZ<F>) * n_sl));
    CK(cudaMalloc(&d_sums, sizeof(XYZZ<F>) * 17 * W));
    CK(cudaMalloc(&d_ref, 8 * 14 * jb.n * W));
    CK(cudaMalloc(&d_got, 8 * 14 * jb.n * W));
    k_fill<<<(total + 127) / 128, 128>>>(d_in, total);
    k_ref<<<1, 32>>>(d_in, jb, W, d_ref);
    CK(cudaGetLastError());
    k_sums_stage1<F><<<n_cta, 128>>>(d_in, jb, W, d_sl);
    CK(cudaGetLastError());
    k_sums_stage2<F><<<W * jb.n, 128>>>(d_sl, jb, W, d_sums);
    CK(cudaGetLastError());
    k_norm<<<1, 32>>>(d_sums, W * jb.n, d_got);
    // direct: the three single elements of the last level, normalised straight from the input array
    uint64_t* d_dir;
    CK(cudaMalloc(&d_dir, 8 * 14 * W));
    k_norm<<<1, 32>>>(d_in + jb.off[3], W, d_dir);
    // and the slice sums of that level as stage 1 left them
    uint64_t* d_slv;
    CK(cudaMalloc(&d_slv, 8 * 14 * W));
    k_norm<<<1, 32>>>(d_sl + jb.sl0[3], W, d_slv);
    CK(cudaDeviceSynchronize());
    std::vector<uint64_t> r(14 * jb.n * W), g(14 * jb.n * W);
    CK(cudaMemcpy(r.data(), d_ref, 8 * r.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g.data(), d_got, 8 * g.size(), cudaMemcpyDeviceToHost));
    run_var<0>(d_in, jb, W, n_cta, d_sl);
    run_var<1>(d_in, jb, W, n_cta, d_sl);
    run_var<2>(d_in, jb, W, n_cta, d_sl);
    std::vector<uint64_t> dd(14 * W), sv(14 * W);
    CK(cudaMemcpy(dd.data(), d_dir, 8 * dd.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(sv.data(), d_slv, 8 * sv.size(), cudaMemcpyDeviceToHost));
    for (uint32_t w = 0; w < W; w++)
        printf("last level w=%u: direct x0=%016llx  stage1 x0=%016llx\n", w, (unsigned long long)dd[w * 14], (unsigned long long)sv[w * 14]);
    int bad = 0;
    for (uint32_t id = 0; id < (uint32_t)jb.n * W; id++) {
        bool ok = true;
        for (int k = 0; k < 13; k++) ok &= r[id * 14 + k] == g[id * 14 + k];
        printf("level %u window %u: %s  ref x0=%016llx flag=%llu | got x0=%016llx flag=%llu\n", id / W, id % W, ok ? "MATCH" : "MISMATCH",
               (unsigned long long)r[id * 14], (unsigned long long)r[id * 14 + 12], (unsigned long long)g[id * 14], (unsigned long long)g[id * 14 + 12]);
        bad += !ok;
    }
    printf("%s\n", bad ? "SUMS CHECK FAILED" : "SUMS CHECK OK");
    return bad != 0;
}
