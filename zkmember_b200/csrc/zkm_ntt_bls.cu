// zkm_ntt_bls.cu -- NTT kernels instantiated for BLS12-381 Fr, plus the curve dispatch.
#include "zkm_ntt.cuh"

namespace zkm {

void ntt_run_bn(Context* c, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int inverse, int coset, cudaStream_t s);
void ntt_domain_constants_bn(Context* c, uint32_t* d, int log_n);
void witness_map_bn(Context* c, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n, uint64_t* d_h, cudaStream_t s);

void fr_into_repr_bn(Context* c, const uint64_t* d_in, uint64_t* d_out, uint64_t n, cudaStream_t s);
void ntt_run_bw6(Context* c, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int inverse, int coset, cudaStream_t s);
void ntt_domain_constants_bw6(Context* c, uint32_t* d, int log_n);
void witness_map_bw6(Context* c, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n, uint64_t* d_h, cudaStream_t s);
void fr_into_repr_bw6(Context* c, const uint64_t* d_in, uint64_t* d_out, uint64_t n, cudaStream_t s);
void kzg_quotient_bn(Context* c, const uint64_t* d_coeffs, size_t n, const uint64_t* d_point, uint64_t* d_quot, uint64_t* d_eval, cudaStream_t s);
void kzg_quotient_bw6(Context* c, const uint64_t* d_coeffs, size_t n, const uint64_t* d_point, uint64_t* d_quot, uint64_t* d_eval, cudaStream_t s);
void kzg_quotient_run(Context* c, int curve, const uint64_t* d_coeffs, size_t n, const uint64_t* d_point, uint64_t* d_quot,
                      uint64_t* d_eval, cudaStream_t s) {
    if (curve == ZKM_CURVE_BLS12_381) kzg_quotient_t<Bls12_381_FrP>(c, d_coeffs, n, d_point, d_quot, d_eval, s);
    else if (curve == ZKM_CURVE_BN254) kzg_quotient_bn(c, d_coeffs, n, d_point, d_quot, d_eval, s);
    else if (curve == ZKM_CURVE_BW6_761) kzg_quotient_bw6(c, d_coeffs, n, d_point, d_quot, d_eval, s);
    else ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
}
void fr_into_repr_run(Context* c, int curve, const uint64_t* d_in, uint64_t* d_out, uint64_t n, cudaStream_t s) {
    if (curve == ZKM_CURVE_BLS12_381) fr_into_repr_t<Bls12_381_FrP>(c, d_in, d_out, n, s);
    else if (curve == ZKM_CURVE_BN254) fr_into_repr_bn(c, d_in, d_out, n, s);
    else if (curve == ZKM_CURVE_BW6_761) fr_into_repr_bw6(c, d_in, d_out, n, s);
    else ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
}

void witness_map_run(Context* c, int curve, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n, uint64_t* d_h,
                     cudaStream_t s) {
    if (curve == ZKM_CURVE_BLS12_381) witness_map_t<Bls12_381_FrP>(c, curve, d_a, d_b, d_c, log_n, d_h, s);
    else if (curve == ZKM_CURVE_BN254) witness_map_bn(c, d_a, d_b, d_c, log_n, d_h, s);
    else if (curve == ZKM_CURVE_BW6_761) witness_map_bw6(c, d_a, d_b, d_c, log_n, d_h, s);
    else ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
}

void ntt_run(Context* c, int curve, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int inverse, int coset,
             cudaStream_t stream) {
    if (curve == ZKM_CURVE_BLS12_381) ntt_run_t<Bls12_381_FrP>(c, curve, d_in, d_out, log_n, inverse, coset, stream);
    else if (curve == ZKM_CURVE_BN254) ntt_run_bn(c, d_in, d_out, log_n, inverse, coset, stream);
    else if (curve == ZKM_CURVE_BW6_761) ntt_run_bw6(c, d_in, d_out, log_n, inverse, coset, stream);
    else ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
}

void ntt_domain_constants(Context* c, int curve, uint32_t log_n, uint64_t* out5x4_host) {
    if (!curve_known(curve)) ZKM_FAIL(ZKM_ERR_ARG, "unknown curve id %d", curve);
    int adicity = fr_two_adicity(curve);
    if ((int)log_n > adicity) ZKM_FAIL(ZKM_ERR_DOMAIN, "log_n %u exceeds the two-adicity %d of Fr", log_n, adicity);
    const size_t bytes = 5 * (size_t)fr_words(curve) * 8;
    uint32_t* d = (uint32_t*)c->io_out.get(bytes);
    if (curve == ZKM_CURVE_BLS12_381) ZKM_LAUNCH(k_domain_constants<Bls12_381_FrP>, 1, 32, 0, c->stream, d, (int)log_n);
    else if (curve == ZKM_CURVE_BN254) ntt_domain_constants_bn(c, d, (int)log_n);
    else ntt_domain_constants_bw6(c, d, (int)log_n);
    ZKM_CUDA(cudaMemcpyAsync(out5x4_host, d, bytes, cudaMemcpyDeviceToHost, c->stream));
    ZKM_CUDA(cudaStreamSynchronize(c->stream));
}

void ntt_release_tables(Shared* sh) {
    for (auto& kv : sh->twiddles) cudaFree(kv.second);
    sh->twiddles.clear();
}

}  // namespace zkm
