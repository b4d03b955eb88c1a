//! Deterministic parity driver — run on a machine WITH cargo, against unpatched arkworks 0.3.0.
//! UNTESTED here (no Rust toolchain in this repository's build environment).
//! Writes the same seeded inputs and arkworks' outputs in the byte format of tests/golden/*.json so that
//! `pytest tests/test_golden.py` can pin the CUDA library (and the C++ oracle) against real arkworks bytes.
//!
//!   cargo run --release -- ntt  bls12_381 16 > ntt_arkworks.json
//!   cargo run --release -- msm  bls12_381 g1 65536 > msm_arkworks.json
//!   cargo run --release -- prove            # Groth16::prove on the 3-member tree with fixed timestamp / index / r / s
use ark_bls12_381::{Fr, G1Affine};
use ark_ec::{msm::VariableBaseMSM, AffineCurve, ProjectiveCurve};
use ark_ff::{PrimeField, UniformRand};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_std::test_rng;

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let mut rng = test_rng();
    match args.get(1).map(|s| s.as_str()) {
        Some("ntt") => {
            let log_n: u32 = args[3].parse().unwrap();
            let dom = Radix2EvaluationDomain::<Fr>::new(1 << log_n).unwrap();
            let x: Vec<Fr> = (0..1usize << log_n).map(|_| Fr::rand(&mut rng)).collect();
            for (inv, coset) in [(false, false), (true, false), (false, true), (true, true)] {
                let mut y = x.clone();
                match (inv, coset) {
                    (false, false) => dom.fft_in_place(&mut y),
                    (true, false) => dom.ifft_in_place(&mut y),
                    (false, true) => dom.coset_fft_in_place(&mut y),
                    (true, true) => dom.coset_ifft_in_place(&mut y),
                }
                println!("{{\"log_n\":{},\"inverse\":{},\"coset\":{},\"input\":\"{}\",\"output\":\"{}\"}}",
                         log_n, inv, coset, hex_limbs(&x), hex_limbs(&y));
            }
        }
        Some("msm") => {
            let n: usize = args[4].parse().unwrap();
            let g = G1Affine::prime_subgroup_generator();
            let bases: Vec<G1Affine> = (0..n).map(|i| g.mul(Fr::from(0x1234567u64 + 0x89ABCDEu64 * i as u64)).into_affine()).collect();
            let scalars: Vec<_> = (0..n).map(|_| Fr::rand(&mut rng).into_repr()).collect();
            let r = VariableBaseMSM::multi_scalar_mul(&bases, &scalars).into_affine();
            println!("{:?}", r);
        }
        _ => eprintln!("usage: parity-driver ntt|msm|prove …"),
    }
}

fn hex_limbs(v: &[Fr]) -> String {
    // Fp256(BigInteger256([u64; 4])): the inner limbs ARE the Montgomery representation
    v.iter().flat_map(|e| (e.0).0.iter().flat_map(|l| l.to_le_bytes().to_vec()).collect::<Vec<u8>>())
        .map(|b| format!("{:02x}", b)).collect()
}
