// Host emulation harness (TEST ONLY): compiles the device headers with
// -DZKM_HOST_EMU so the exact limb schedules of the kernels run on the CPU and can
// be compared with Python big integers through ctypes.  Not part of the product.
#include <cstring>
#include "../../zkmember_b200/csrc/zkm_curve.cuh"

using namespace zkm;

template <class F>
static void fp_ops(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int count) {
    constexpr int N = F::N;
    for (int k = 0; k < count; k++) {
        F x, y, r;
        memcpy(x.l, a + k * N, 4 * N);
        memcpy(y.l, b + k * N, 4 * N);
        switch (op) {
            case 0: r = x * y; break;
            case 1: r = x + y; break;
            case 2: r = x - y; break;
            case 3: r = neg(x); break;
            case 4: r = sqr(x); break;
            case 5: r = inv(x); break;
            case 6: r = dbl(x); break;
            case 7: r = fp_inv_fermat(x); break;
            case 8: r = fp_mul_unsat(x, y); break;   // unsaturated-radix column product (zkm_fpmul_u.cuh)
            case 9: r = fp_sqr_unsat(x); break;
            default: r = F::zero();
        }
        memcpy(out + k * N, r.l, 4 * N);
    }
}

template <class P>
static void fp2_ops(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int count) {
    constexpr int N = P::N;
    typedef Fp2<P> F2;
    for (int k = 0; k < count; k++) {
        F2 x, y, r;
        memcpy(x.c0.l, a + k * 2 * N, 4 * N);
        memcpy(x.c1.l, a + k * 2 * N + N, 4 * N);
        memcpy(y.c0.l, b + k * 2 * N, 4 * N);
        memcpy(y.c1.l, b + k * 2 * N + N, 4 * N);
        switch (op) {
            case 0: r = x * y; break;
            case 1: r = x + y; break;
            case 2: r = x - y; break;
            case 3: r = neg(x); break;
            case 4: r = sqr(x); break;
            case 5: r = inv(x); break;
            case 6: r = dbl(x); break;
            default: r = F2::zero();
        }
        memcpy(out + k * 2 * N, r.c0.l, 4 * N);
        memcpy(out + k * 2 * N + N, r.c1.l, 4 * N);
    }
}

// XYZZ buffers: 4 coordinates; affine buffers: 2 coordinates (+ separate flag)
template <class F>
static void curve_ops(int op, const uint32_t* p, const uint32_t* q, uint32_t* out, uint8_t* out_inf, int count) {
    constexpr int W = sizeof(F) / 4;
    for (int k = 0; k < count; k++) {
        XYZZ<F> P;
        memcpy(&P, p + k * 4 * W, 16 * W);
        switch (op) {
            case 0: {  // madd: q affine
                F x, y;
                memcpy(&x, q + k * 2 * W, 4 * W);
                memcpy(&y, q + k * 2 * W + W, 4 * W);
                xyzz_madd(P, x, y);
                memcpy(out + k * 4 * W, &P, 16 * W);
                break;
            }
            case 1: {  // add: q xyzz
                XYZZ<F> Q;
                memcpy(&Q, q + k * 4 * W, 16 * W);
                xyzz_add(P, Q);
                memcpy(out + k * 4 * W, &P, 16 * W);
                break;
            }
            case 2: {  // dbl
                xyzz_dbl(P);
                memcpy(out + k * 4 * W, &P, 16 * W);
                break;
            }
            case 3: {  // to_affine: out 2 coords + flag
                F x = F::zero(), y = F::zero();
                bool ok = xyzz_to_affine(P, x, y);
                memcpy(out + k * 2 * W, &x, 4 * W);
                memcpy(out + k * 2 * W + W, &y, 4 * W);
                out_inf[k] = ok ? 0 : 1;
                break;
            }
        }
    }
}

extern "C" {
// field: 0 bls fq, 1 bls fr, 2 bn fq, 3 bn fr, 4 bw6 fq (24 limbs), 5 bw6 fr (12 limbs)
int emu_fp_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int count) {
    switch (field) {
        case 0: fp_ops<Bls12_381_Fq>(op, a, b, out, count); return 0;
        case 1: fp_ops<Bls12_381_Fr>(op, a, b, out, count); return 0;
        case 2: fp_ops<Bn254_Fq>(op, a, b, out, count); return 0;
        case 3: fp_ops<Bn254_Fr>(op, a, b, out, count); return 0;
        case 4: fp_ops<Bw6_761_Fq>(op, a, b, out, count); return 0;
        case 5: fp_ops<Bw6_761_Fr>(op, a, b, out, count); return 0;
    }
    return -1;
}
int emu_fp2_op(int curve, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int count) {
    if (curve == 0) fp2_ops<Bls12_381_FqP>(op, a, b, out, count);
    else if (curve == 1) fp2_ops<Bn254_FqP>(op, a, b, out, count);
    else return -1;
    return 0;
}
int emu_curve_op(int curve, int group, int op, const uint32_t* p, const uint32_t* q, uint32_t* out, uint8_t* out_inf, int count) {
    if (curve == 0 && group == 1) curve_ops<Bls12_381_Fq>(op, p, q, out, out_inf, count);
    else if (curve == 0 && group == 2) curve_ops<Bls12_381_Fq2>(op, p, q, out, out_inf, count);
    else if (curve == 1 && group == 1) curve_ops<Bn254_Fq>(op, p, q, out, out_inf, count);
    else if (curve == 1 && group == 2) curve_ops<Bn254_Fq2>(op, p, q, out, out_inf, count);
    else if (curve == 2) curve_ops<Bw6_761_Fq>(op, p, q, out, out_inf, count);   // G1 and G2 both over Fq
    else return -1;
    return 0;
}
}
