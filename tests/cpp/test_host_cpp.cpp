// C++ host-layer parity test: drives include/zkm_b200.hpp (the compiled-language mirror of the arkworks
// interfaces) over known-answer vectors handed in as text lines by tests/test_host_cpp.py:
//   ntt <curve> <log_n> <inverse> <coset> <hex input> <hex output>
//   msm <curve> <group> <n> <hex bases> <hex infinity flags> <hex scalars> <hex result> <result infinity>
//   wmap <curve> <log_n> <hex a> <hex b> <hex c> <hex h>
//   ser <curve> <group> <hex xy> <infinity> <hex compressed>     (host-only: checked BEFORE zkm::init, no GPU needed)
//   kzg <curve> <n> <hex powers> <hex gamma powers> <hex coeffs> <hex blinding> <hex z> then four (hex xy, infinity) pairs --
//       commit, hiding commit, open.w, hiding open.w -- and <hex random_v>
// Byte equality is the bar.  Exit code 0 = all vectors match.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/zkm_b200.hpp"

static std::vector<uint8_t> unhex(const std::string& s) {
    std::vector<uint8_t> out;
    if (s == "-") return out;
    out.reserve(s.size() / 2);
    auto v = [](char c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; };
    for (size_t i = 0; i + 1 < s.size(); i += 2) out.push_back((uint8_t)(v(s[i]) * 16 + v(s[i + 1])));
    return out;
}

template <class Curve>
static bool run_ntt(int log_n, int inverse, int coset, const std::vector<uint8_t>& in, const std::vector<uint8_t>& want) {
    typedef typename Curve::Fr F;
    auto dom = zkm::Radix2EvaluationDomain<Curve>::new_(size_t(1) << log_n);
    if (!dom) return false;
    std::vector<F> v(in.size() / sizeof(F));
    std::memcpy(v.data(), in.data(), in.size());
    std::vector<F> r = inverse ? (coset ? dom->coset_ifft(v) : dom->ifft(v)) : (coset ? dom->coset_fft(v) : dom->fft(v));
    return r.size() * sizeof(F) == want.size() && std::memcmp(r.data(), want.data(), want.size()) == 0;
}

template <class Curve, int GROUP>
static bool run_msm(size_t n, const std::vector<uint8_t>& bases, const std::vector<uint8_t>& inf, const std::vector<uint8_t>& scal,
                    const std::vector<uint8_t>& want, int want_inf) {
    typedef zkm::GroupAffine<Curve, GROUP> A;
    typedef typename A::Coord Coord;
    std::vector<A> b(n);
    for (size_t i = 0; i < n; i++) {
        std::memcpy(&b[i].x, bases.data() + i * 2 * sizeof(Coord), sizeof(Coord));
        std::memcpy(&b[i].y, bases.data() + i * 2 * sizeof(Coord) + sizeof(Coord), sizeof(Coord));
        b[i].infinity = inf[i] != 0;
    }
    std::vector<typename Curve::BigInt> s(n);
    if (n) std::memcpy(s.data(), scal.data(), n * sizeof(typename Curve::BigInt));
    A r = zkm::VariableBaseMSM::multi_scalar_mul<Curve, GROUP>(b, s);
    bool ok = (int)r.infinity == want_inf && std::memcmp(&r.x, want.data(), sizeof(Coord)) == 0 &&
              std::memcmp(&r.y, want.data() + sizeof(Coord), sizeof(Coord)) == 0;
    // the registered path must agree (plain and with precomputed window multiples)
    for (int pre = 0; pre < 2 && ok && n; pre++) {
        zkm::RegisteredBases<Curve, GROUP> reg(b, pre != 0);
        ok = reg.msm(s) == r;
    }
    return ok;
}

template <class Curve>
static bool run_wmap(int log_n, const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, const std::vector<uint8_t>& c,
                     const std::vector<uint8_t>& want) {
    typedef typename Curve::Fr F;
    auto dom = zkm::Radix2EvaluationDomain<Curve>::new_(size_t(1) << log_n);
    size_t n = size_t(1) << log_n;
    std::vector<F> va(n), vb(n), vc(n);
    std::memcpy(va.data(), a.data(), n * sizeof(F));
    std::memcpy(vb.data(), b.data(), n * sizeof(F));
    std::memcpy(vc.data(), c.data(), n * sizeof(F));
    std::vector<F> h = zkm::witness_map(*dom, va, vb, vc);
    return std::memcmp(h.data(), want.data(), n * sizeof(F)) == 0;
}

template <class Curve>
static bool run_kzg(size_t n, const std::vector<uint8_t>& pw, const std::vector<uint8_t>& gpw, const std::vector<uint8_t>& co,
                    const std::vector<uint8_t>& bl, const std::vector<uint8_t>& z, const std::vector<std::vector<uint8_t>>& want,
                    const int* want_inf, const std::vector<uint8_t>& want_rv) {
    typedef zkm::GroupAffine<Curve, 1> A;
    typedef typename A::Coord Coord;
    typedef typename Curve::Fr F;
    auto points = [&](const std::vector<uint8_t>& raw) {
        std::vector<A> b(n);
        for (size_t i = 0; i < n; i++) {
            std::memcpy(&b[i].x, raw.data() + i * 2 * sizeof(Coord), sizeof(Coord));
            std::memcpy(&b[i].y, raw.data() + i * 2 * sizeof(Coord) + sizeof(Coord), sizeof(Coord));
            b[i].infinity = false;
        }
        return b;
    };
    auto elems = [&](const std::vector<uint8_t>& raw) {
        std::vector<F> v(raw.size() / sizeof(F));
        std::memcpy(v.data(), raw.data(), raw.size());
        return v;
    };
    auto same = [&](const A& p, int k) {
        return (int)p.infinity == want_inf[k] && std::memcmp(&p.x, want[k].data(), sizeof(Coord)) == 0 &&
               std::memcmp(&p.y, want[k].data() + sizeof(Coord), sizeof(Coord)) == 0;
    };
    zkm::RegisteredBases<Curve, 1> g(points(pw)), gg(points(gpw), true);
    std::vector<F> c = elems(co), b = elems(bl);
    F point = elems(z)[0];
    bool ok = same(zkm::KZG10::commit<Curve>(g, c), 0) && same(zkm::KZG10::commit<Curve>(g, gg, c, b), 1);
    auto batch = zkm::KZG10::commit_batch<Curve>(g, {c, b, c});
    ok = ok && batch.size() == 3 && same(batch[0], 0) && same(batch[2], 0);
    auto pr = zkm::KZG10::open<Curve>(g, c, point);
    ok = ok && same(pr.w, 2) && !pr.hiding;
    auto prh = zkm::KZG10::open<Curve>(g, c, point, &gg, &b);
    ok = ok && same(prh.w, 3) && prh.hiding && std::memcmp(&prh.random_v, want_rv.data(), sizeof(F)) == 0;
    return ok;
}

template <class Curve, int GROUP>
static bool run_ser(const std::vector<uint8_t>& xy, int inf, const std::vector<uint8_t>& want) {
    typedef zkm::GroupAffine<Curve, GROUP> A;
    typedef typename A::Coord Coord;
    A pt;
    std::memcpy(&pt.x, xy.data(), sizeof(Coord));
    std::memcpy(&pt.y, xy.data() + sizeof(Coord), sizeof(Coord));
    pt.infinity = inf != 0;
    return zkm::serialize(pt) == want;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s vectors.txt\n", argv[0]); return 2; }
    {   // arkworks compressed serialization: pure host code, runs with or without a GPU
        std::ifstream f(argv[1]);
        std::string line;
        int total = 0, bad = 0;
        while (std::getline(f, line)) {
            std::istringstream is(line);
            std::string kind, curve, hxy, hwant;
            int group, inf;
            is >> kind;
            if (kind != "ser") continue;
            is >> curve >> group >> hxy >> inf >> hwant;
            auto XY = unhex(hxy), W = unhex(hwant);
            bool ok;
            if (curve == "bls12_381") ok = group == 1 ? run_ser<zkm::Bls12_381, 1>(XY, inf, W) : run_ser<zkm::Bls12_381, 2>(XY, inf, W);
            else if (curve == "bn254") ok = group == 1 ? run_ser<zkm::Bn254, 1>(XY, inf, W) : run_ser<zkm::Bn254, 2>(XY, inf, W);
            else ok = group == 1 ? run_ser<zkm::Bw6_761, 1>(XY, inf, W) : run_ser<zkm::Bw6_761, 2>(XY, inf, W);
            total++;
            if (!ok) { bad++; std::printf("MISMATCH: ser %s g%d (vector %d)\n", curve.c_str(), group, total); }
        }
        std::printf("ser: %d vectors, %d mismatches\n", total, bad);
        if (bad) return 1;
    }
    try {
        zkm::init(0);
    } catch (const zkm::Error& e) {
        std::printf("INIT FAILED (%d): %s\n", e.code, e.what());   // expected on a box without a GPU: no CPU fallback
        return 3;
    }
    // error behaviour mirrors upstream: no domain above the two-adicity
    if (zkm::Radix2EvaluationDomain<zkm::Bn254>::new_((size_t(1) << 28) + 1)) { std::printf("FAIL: domain above two-adicity\n"); return 1; }
    std::ifstream f(argv[1]);
    std::string line;
    int total = 0, bad = 0;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        std::string kind, curve;
        is >> kind >> curve;
        bool ok = false;
        try {
            if (kind == "ntt") {
                int log_n, inv, cos; std::string hin, hout;
                is >> log_n >> inv >> cos >> hin >> hout;
                ok = curve == "bls12_381" ? run_ntt<zkm::Bls12_381>(log_n, inv, cos, unhex(hin), unhex(hout))
                     : curve == "bn254"   ? run_ntt<zkm::Bn254>(log_n, inv, cos, unhex(hin), unhex(hout))
                                          : run_ntt<zkm::Bw6_761>(log_n, inv, cos, unhex(hin), unhex(hout));
            } else if (kind == "msm") {
                int group, rinf; size_t n; std::string hb, hi, hs, hr;
                is >> group >> n >> hb >> hi >> hs >> hr >> rinf;
                auto B = unhex(hb), I = unhex(hi), S = unhex(hs), R = unhex(hr);
                if (curve == "bls12_381") ok = group == 1 ? run_msm<zkm::Bls12_381, 1>(n, B, I, S, R, rinf) : run_msm<zkm::Bls12_381, 2>(n, B, I, S, R, rinf);
                else if (curve == "bn254") ok = group == 1 ? run_msm<zkm::Bn254, 1>(n, B, I, S, R, rinf) : run_msm<zkm::Bn254, 2>(n, B, I, S, R, rinf);
                else ok = group == 1 ? run_msm<zkm::Bw6_761, 1>(n, B, I, S, R, rinf) : run_msm<zkm::Bw6_761, 2>(n, B, I, S, R, rinf);
            } else if (kind == "wmap") {
                int log_n; std::string ha, hb, hc, hh;
                is >> log_n >> ha >> hb >> hc >> hh;
                ok = curve == "bls12_381" ? run_wmap<zkm::Bls12_381>(log_n, unhex(ha), unhex(hb), unhex(hc), unhex(hh))
                     : curve == "bn254"   ? run_wmap<zkm::Bn254>(log_n, unhex(ha), unhex(hb), unhex(hc), unhex(hh))
                                          : run_wmap<zkm::Bw6_761>(log_n, unhex(ha), unhex(hb), unhex(hc), unhex(hh));
            } else if (kind == "kzg") {
                size_t n; std::string hp, hg, hc, hb, hz, hrv;
                std::vector<std::vector<uint8_t>> want(4);
                int winf[4];
                is >> n >> hp >> hg >> hc >> hb >> hz;
                for (int k = 0; k < 4; k++) { std::string h; is >> h >> winf[k]; want[k] = unhex(h); }
                is >> hrv;
                ok = curve == "bls12_381" ? run_kzg<zkm::Bls12_381>(n, unhex(hp), unhex(hg), unhex(hc), unhex(hb), unhex(hz), want, winf, unhex(hrv))
                     : curve == "bn254"   ? run_kzg<zkm::Bn254>(n, unhex(hp), unhex(hg), unhex(hc), unhex(hb), unhex(hz), want, winf, unhex(hrv))
                                          : run_kzg<zkm::Bw6_761>(n, unhex(hp), unhex(hg), unhex(hc), unhex(hb), unhex(hz), want, winf, unhex(hrv));
            } else {
                continue;
            }
        } catch (const zkm::Error& e) {
            std::printf("ERROR (%d): %s\n", e.code, e.what());
        }
        total++;
        if (!ok) { bad++; std::printf("MISMATCH: %s %s (vector %d)\n", kind.c_str(), curve.c_str(), total); }
    }
    std::printf("host_cpp: %d vectors, %d mismatches\n", total, bad);
    zkm_shutdown();
    return bad ? 1 : (total ? 0 : 2);
}
