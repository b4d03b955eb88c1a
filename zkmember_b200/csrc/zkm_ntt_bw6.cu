// zkm_ntt_bw6.cu -- NTT kernels instantiated for the 377-bit Fr of BW6-761 (= Fq of BLS12-377, 12 x 32-bit limbs).
#include "zkm_ntt.cuh"

namespace zkm {

void ntt_run_bw6(Context* c, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int inverse, int coset, cudaStream_t s) {
    ntt_run_t<Bw6_761_FrP>(c, ZKM_CURVE_BW6_761, d_in, d_out, log_n, inverse, coset, s);
}
void witness_map_bw6(Context* c, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n, uint64_t* d_h, cudaStream_t s) {
    witness_map_t<Bw6_761_FrP>(c, ZKM_CURVE_BW6_761, d_a, d_b, d_c, log_n, d_h, s);
}
void fr_into_repr_bw6(Context* c, const uint64_t* d_in, uint64_t* d_out, uint64_t n, cudaStream_t s) {
    fr_into_repr_t<Bw6_761_FrP>(c, d_in, d_out, n, s);
}
void ntt_domain_constants_bw6(Context* c, uint32_t* d, int log_n) {
    ZKM_LAUNCH(k_domain_constants<Bw6_761_FrP>, 1, 32, 0, c->stream, d, log_n);
}
void kzg_quotient_bw6(Context* c, const uint64_t* d_coeffs, size_t n, const uint64_t* d_point, uint64_t* d_quot, uint64_t* d_eval,
                     cudaStream_t s) {
    kzg_quotient_t<Bw6_761_FrP>(c, d_coeffs, n, d_point, d_quot, d_eval, s);
}

}  // namespace zkm
