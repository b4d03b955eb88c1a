// zkm_b200.hpp -- header-only C++ host layer above the C ABI (include/zkm_b200.h).
//
// zkMember is compiled Rust and no Rust toolchain exists in this build environment, so the compiled-language
// host side of the drop-in is this C++ mirror of the two arkworks interfaces the hot path sits behind
// (same names, argument meaning and error behaviour); INTEGRATION.md shows the equivalent Rust shim.
//
//   zkm::VariableBaseMSM::multi_scalar_mul(bases, scalars)      ark-ec 0.3.0 src/msm/variable_base.rs
//   zkm::Radix2EvaluationDomain<Curve>::new_(n) / fft / ifft / coset_fft / coset_ifft (+ _in_place)
//                                                               ark-poly 0.3.0 src/domain/radix2/mod.rs
//   zkm::witness_map(domain, a, b, c)                           ark-groth16 0.3.0 src/r1cs_to_qap.rs
//   zkm::ProvingKeyBases / KZG10::commit                        registered bases (pk / SRS reuse)
//
// Types are plain structs with arkworks' memory layout: Fp256 = 4 x u64 Montgomery limbs, Fp384 = 6,
// BigInteger256 = 4 x u64 canonical limbs (BigInteger384 = 6 for BW6-761), GroupAffine{x, y, infinity}.  Errors of the C ABI become
// zkm::Error exceptions (upstream's functions are infallible; there is no CPU fallback to hide behind).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "zkm_b200.h"

namespace zkm {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int32_t rc, const char* what) {
    if (rc != ZKM_OK) throw Error(rc, std::string(what) + ": " + zkm_last_error());
}
inline void init(int device = 0) { check(zkm_init(device), "zkm_init"); }

template <int L64>
struct Fp {
    uint64_t limbs[L64];  // Montgomery form, little-endian limbs (ark-ff Fp256 / Fp384)
    bool operator==(const Fp& o) const { return std::memcmp(limbs, o.limbs, sizeof(limbs)) == 0; }
};
template <int L64>
struct BigInteger {
    uint64_t limbs[L64];  // canonical integer (Fr::into_repr())
};
typedef BigInteger<4> BigInteger256;
typedef BigInteger<6> BigInteger384;   // scalars of BW6-761 (377-bit Fr)
template <int L64>
struct Fp2 {
    Fp<L64> c0, c1;
};

struct Bls12_381 {
    static constexpr int ID = ZKM_CURVE_BLS12_381;
    static constexpr int TWO_ADICITY = 32;
    typedef Fp<4> Fr;
    typedef Fp<6> Fq;
    typedef Fp2<6> G2Coord;
    typedef BigInteger256 BigInt;
};
struct Bn254 {
    static constexpr int ID = ZKM_CURVE_BN254;
    static constexpr int TWO_ADICITY = 28;
    typedef Fp<4> Fr;
    typedef Fp<4> Fq;
    typedef Fp2<4> G2Coord;
    typedef BigInteger256 BigInt;
};
// The second curve zkMember instantiates (/root/reference/benches/groth16.rs:24-29): 761-bit Fq, 377-bit Fr
// (= the base field of BLS12-377), and a G2 that is a curve over Fq itself.
struct Bw6_761 {
    static constexpr int ID = ZKM_CURVE_BW6_761;
    static constexpr int TWO_ADICITY = 46;
    typedef Fp<6> Fr;
    typedef Fp<12> Fq;
    typedef Fp<12> G2Coord;
    typedef BigInteger384 BigInt;
};

// GroupAffine<P>: x, y, infinity.  G1: Coord = Fq; G2: Coord = Fp2<Fq limbs> (BW6-761: Fq).
template <class Curve, int GROUP>
struct GroupAffine {
    typedef typename std::conditional<GROUP == 1, typename Curve::Fq, typename Curve::G2Coord>::type Coord;
    Coord x, y;
    bool infinity = false;
    bool operator==(const GroupAffine& o) const {
        return infinity == o.infinity && std::memcmp(&x, &o.x, sizeof(Coord)) == 0 && std::memcmp(&y, &o.y, sizeof(Coord)) == 0;
    }
};
template <class Curve> using G1Affine = GroupAffine<Curve, 1>;
template <class Curve> using G2Affine = GroupAffine<Curve, 2>;

namespace detail {
template <class Curve, int GROUP>
inline void pack(const std::vector<GroupAffine<Curve, GROUP>>& bases, size_t n, std::vector<uint64_t>& xy, std::vector<uint8_t>& inf) {
    typedef typename GroupAffine<Curve, GROUP>::Coord Coord;
    constexpr size_t W = sizeof(Coord) / 8;
    xy.resize(n * 2 * W);
    inf.resize(n);
    for (size_t i = 0; i < n; i++) {   // GroupAffine is not guaranteed packed: copy field by field
        std::memcpy(&xy[i * 2 * W], &bases[i].x, sizeof(Coord));
        std::memcpy(&xy[i * 2 * W + W], &bases[i].y, sizeof(Coord));
        inf[i] = bases[i].infinity ? 1 : 0;
    }
}
template <class Curve, int GROUP>
inline GroupAffine<Curve, GROUP> unpack(const std::vector<uint64_t>& out, uint8_t out_inf) {
    typedef typename GroupAffine<Curve, GROUP>::Coord Coord;
    constexpr size_t W = sizeof(Coord) / 8;
    GroupAffine<Curve, GROUP> r;
    std::memcpy(&r.x, &out[0], sizeof(Coord));
    std::memcpy(&r.y, &out[W], sizeof(Coord));
    r.infinity = out_inf != 0;
    return r;
}
}  // namespace detail

// Bases uploaded once (proving-key query vectors, KZG powers); `precompute` stores the window multiples.
template <class Curve, int GROUP>
class RegisteredBases {
public:
    RegisteredBases(const std::vector<GroupAffine<Curve, GROUP>>& bases, bool precompute = false) : n_(bases.size()) {
        std::vector<uint64_t> xy;
        std::vector<uint8_t> inf;
        detail::pack<Curve, GROUP>(bases, n_, xy, inf);
        if (precompute) check(zkm_set_option("msm_precompute", 1), "zkm_set_option");
        int32_t rc = zkm_bases_register(Curve::ID, GROUP, xy.data(), inf.data(), n_, &handle_);
        if (precompute) zkm_set_option("msm_precompute", 0);
        check(rc, "zkm_bases_register");
    }
    ~RegisteredBases() { if (handle_) zkm_bases_release(handle_); }
    RegisteredBases(const RegisteredBases&) = delete;
    RegisteredBases& operator=(const RegisteredBases&) = delete;
    size_t size() const { return n_; }
    uint64_t handle() const { return handle_; }
    // sum over min(size() - offset, scalars.size()) pairs starting at bases[offset]
    GroupAffine<Curve, GROUP> msm(const std::vector<typename Curve::BigInt>& scalars, size_t offset = 0) const {
        size_t n = std::min(scalars.size(), n_ - std::min(offset, n_));
        typedef typename GroupAffine<Curve, GROUP>::Coord Coord;
        std::vector<uint64_t> out(2 * sizeof(Coord) / 8);
        uint8_t out_inf = 0;
        check(zkm_msm_registered(handle_, offset, reinterpret_cast<const uint64_t*>(scalars.data()), n, out.data(), &out_inf),
              "zkm_msm_registered");
        return detail::unpack<Curve, GROUP>(out, out_inf);
    }

private:
    size_t n_;
    uint64_t handle_ = 0;
};

struct VariableBaseMSM {
    // ark_ec::msm::VariableBaseMSM::multi_scalar_mul followed by into_affine()
    template <class Curve, int GROUP>
    static GroupAffine<Curve, GROUP> multi_scalar_mul(const std::vector<GroupAffine<Curve, GROUP>>& bases,
                                                      const std::vector<typename Curve::BigInt>& scalars) {
        typedef typename GroupAffine<Curve, GROUP>::Coord Coord;
        const size_t size = std::min(bases.size(), scalars.size());   // upstream: min(bases.len(), scalars.len())
        std::vector<uint64_t> xy;
        std::vector<uint8_t> inf;
        detail::pack<Curve, GROUP>(bases, size, xy, inf);
        std::vector<uint64_t> out(2 * sizeof(Coord) / 8);
        uint8_t out_inf = 0;
        const uint64_t* sc = reinterpret_cast<const uint64_t*>(scalars.data());
        int32_t rc = GROUP == 1 ? zkm_msm_g1(Curve::ID, xy.data(), inf.data(), sc, size, out.data(), &out_inf)
                                : zkm_msm_g2(Curve::ID, xy.data(), inf.data(), sc, size, out.data(), &out_inf);
        check(rc, "zkm_msm");
        return detail::unpack<Curve, GROUP>(out, out_inf);
    }
};

template <class Curve>
class Radix2EvaluationDomain {
public:
    typedef typename Curve::Fr F;
    uint64_t size;
    uint32_t log_size_of_group;
    F size_inv, group_gen, group_gen_inv, generator_inv;

    // Radix2EvaluationDomain::new: None when the field has no subgroup of that order
    static std::optional<Radix2EvaluationDomain> new_(size_t num_coeffs) {
        size_t n = num_coeffs ? num_coeffs : 1;
        uint32_t log_n = 0;
        while ((size_t(1) << log_n) < n) log_n++;
        if ((int)log_n > Curve::TWO_ADICITY) return std::nullopt;
        Radix2EvaluationDomain d;
        d.size = uint64_t(1) << log_n;
        d.log_size_of_group = log_n;
        constexpr size_t S = sizeof(F) / 8;
        uint64_t c[5 * S];
        check(zkm_domain_constants(Curve::ID, log_n, c), "zkm_domain_constants");
        std::memcpy(&d.group_gen, c, sizeof(F));
        std::memcpy(&d.group_gen_inv, c + S, sizeof(F));
        std::memcpy(&d.size_inv, c + 2 * S, sizeof(F));
        std::memcpy(&d.generator_inv, c + 4 * S, sizeof(F));
        return d;
    }
    void fft_in_place(std::vector<F>& coeffs) const { run(coeffs, 0, 0); }
    void ifft_in_place(std::vector<F>& evals) const { run(evals, 1, 0); }
    void coset_fft_in_place(std::vector<F>& coeffs) const { run(coeffs, 0, 1); }
    void coset_ifft_in_place(std::vector<F>& evals) const { run(evals, 1, 1); }
    std::vector<F> fft(std::vector<F> v) const { fft_in_place(v); return v; }
    std::vector<F> ifft(std::vector<F> v) const { ifft_in_place(v); return v; }
    std::vector<F> coset_fft(std::vector<F> v) const { coset_fft_in_place(v); return v; }
    std::vector<F> coset_ifft(std::vector<F> v) const { coset_ifft_in_place(v); return v; }

private:
    void run(std::vector<F>& v, int inverse, int coset) const {
        if (v.size() > size) throw Error(ZKM_ERR_ARG, "input longer than the domain");
        F zero;
        std::memset(&zero, 0, sizeof(zero));
        v.resize(size, zero);   // upstream: coeffs.resize(self.size(), T::zero())
        check(zkm_ntt(Curve::ID, reinterpret_cast<uint64_t*>(v.data()), log_size_of_group, inverse, coset), "zkm_ntt");
    }
};

// ark_groth16::R1CStoQAP::witness_map, the part after the matrix-vector products
template <class Curve>
inline std::vector<typename Curve::Fr> witness_map(const Radix2EvaluationDomain<Curve>& d, const std::vector<typename Curve::Fr>& a,
                                                   const std::vector<typename Curve::Fr>& b, const std::vector<typename Curve::Fr>& c) {
    if (a.size() != d.size || b.size() != d.size || c.size() != d.size) throw Error(ZKM_ERR_ARG, "a, b, c must have the domain size");
    std::vector<typename Curve::Fr> h(d.size);
    check(zkm_witness_map(Curve::ID, reinterpret_cast<const uint64_t*>(a.data()), reinterpret_cast<const uint64_t*>(b.data()),
                          reinterpret_cast<const uint64_t*>(c.data()), d.log_size_of_group, reinterpret_cast<uint64_t*>(h.data())),
          "zkm_witness_map");
    return h;
}

struct KZG10 {
    // non-hiding part of ark_poly_commit::kzg10::KZG10::commit
    template <class Curve>
    static G1Affine<Curve> commit(const RegisteredBases<Curve, 1>& powers_of_g, const std::vector<typename Curve::Fr>& coeffs) {
        std::vector<uint64_t> out(2 * sizeof(typename Curve::Fq) / 8);
        uint8_t out_inf = 0;
        check(zkm_kzg_commit(powers_of_g.handle(), reinterpret_cast<const uint64_t*>(coeffs.data()), coeffs.size(), out.data(), &out_inf),
              "zkm_kzg_commit");
        return detail::unpack<Curve, 1>(out, out_inf);
    }
};

}  // namespace zkm
