//! Deterministic parity driver -- run on a machine WITH cargo, against UNPATCHED arkworks 0.3.0 (the versions pinned in
//! /root/reference/Cargo.lock).  UNCOMPILED here (no Rust toolchain in this repository's build environment).
//! Prints arkworks' own inputs and outputs in the byte format of tests/golden/*.json, so that
//! `pytest tests/test_golden.py` can pin the CUDA library AND the C++ / Python oracles against real arkworks bytes
//! (today they are pinned against exact big-int definitions only -- "parity unpinned").
//!
//!   cargo run --release -- ntt   <log_n>            > ntt_arkworks.json      four transforms of one seeded vector
//!   cargo run --release -- msm   <n>                > msm_arkworks.json      G1 MSM, bases (a0 + i d) G, seeded scalars
//!   cargo run --release -- prove                    > proof_arkworks.json    Groth16 on zkMember's 3-member tree with a
//!                                                                            FIXED join date, index and (r, s)
//! Cargo.toml of this crate: ark-bls12-381, ark-ec, ark-ff, ark-poly, ark-groth16, ark-relations, ark-serialize,
//! ark-std (all 0.3), ark-crypto-primitives 0.3 (r1cs), and `zkmember = { path = "<reference>" }`.
use ark_bls12_381::{Bls12_381, Fr, G1Affine};
use ark_ec::{msm::VariableBaseMSM, AffineCurve, ProjectiveCurve};
use ark_ff::{PrimeField, UniformRand};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::CanonicalSerialize;
use ark_std::test_rng;

fn hex(bytes: &[u8]) -> String { bytes.iter().map(|b| format!("{:02x}", b)).collect() }

/// Montgomery limbs of Fr / Fq elements: `Fp256(pub BigInteger256([u64; 4]), _)` -- the inner limbs ARE the
/// representation the C ABI exchanges.
fn fr_limbs(v: &[Fr]) -> String {
    hex(&v.iter().flat_map(|e| (e.0).0.iter().flat_map(|l| l.to_le_bytes().to_vec()).collect::<Vec<u8>>()).collect::<Vec<u8>>())
}
fn g1_xy(p: &G1Affine) -> String {
    let mut out = Vec::new();
    for c in [&p.x, &p.y] { for l in (c.0).0.iter() { out.extend_from_slice(&l.to_le_bytes()); } }
    hex(&out)
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let mut rng = test_rng();
    match args.get(1).map(|s| s.as_str()) {
        Some("ntt") => {
            let log_n: u32 = args[2].parse().unwrap();
            let dom = Radix2EvaluationDomain::<Fr>::new(1 << log_n).unwrap();
            let x: Vec<Fr> = (0..1usize << log_n).map(|_| Fr::rand(&mut rng)).collect();
            let mut rows = Vec::new();
            for (inv, coset) in [(false, false), (true, false), (false, true), (true, true)] {
                let mut y = x.clone();
                match (inv, coset) {
                    (false, false) => dom.fft_in_place(&mut y),
                    (true, false) => dom.ifft_in_place(&mut y),
                    (false, true) => dom.coset_fft_in_place(&mut y),
                    (true, true) => dom.coset_ifft_in_place(&mut y),
                }
                rows.push(format!("{{\"curve\":\"bls12_381\",\"log_n\":{},\"inverse\":{},\"coset\":{},\"input\":\"{}\",\"output\":\"{}\"}}",
                                  log_n, inv, coset, fr_limbs(&x), fr_limbs(&y)));
            }
            println!("{{\"generator\":\"rust/parity-driver (arkworks 0.3.0)\",\"vectors\":[{}]}}", rows.join(","));
        }
        Some("msm") => {
            let n: usize = args[2].parse().unwrap();
            let g = G1Affine::prime_subgroup_generator();
            let bases: Vec<G1Affine> =
                (0..n).map(|i| g.mul(Fr::from(0x1234567u64 + 0x89ABCDEu64 * i as u64)).into_affine()).collect();
            let scalars: Vec<_> = (0..n).map(|_| Fr::rand(&mut rng).into_repr()).collect();
            let r = VariableBaseMSM::multi_scalar_mul(&bases, &scalars).into_affine();
            let sc: Vec<u8> = scalars.iter().flat_map(|s| s.0.iter().flat_map(|l| l.to_le_bytes().to_vec()).collect::<Vec<u8>>()).collect();
            let zero_xy = "00".repeat(96);
            println!("{{\"generator\":\"rust/parity-driver (arkworks 0.3.0)\",\"vectors\":[{{\"curve\":\"bls12_381\",\"group\":1,\"n\":{},\
                      \"bases\":\"{}\",\"infinity\":[{}],\"scalars\":\"{}\",\"result\":\"{}\",\"result_infinity\":{}}}]}}",
                     n, bases.iter().map(g1_xy).collect::<String>(), vec!["0"; n].join(","), hex(&sc),
                     if r.infinity { zero_xy } else { g1_xy(&r) }, r.infinity as u8);
        }
        Some("prove") => prove(),
        _ => eprintln!("usage: parity-driver ntt <log_n> | msm <n> | prove"),
    }
}

/// Groth16 over zkMember's own circuit with every source of nondeterminism pinned (the bench stamps `Utc::now()` into
/// each member -- /root/reference/src/member.rs:28,40 -- and draws the index from an unseeded RNG,
/// benches/groth16.rs:94): fixed join date, index 1, `r`, `s` from `test_rng()`.  Output: the evaluation vectors are
/// not exposed by ark-groth16, so the record holds what a parity test needs end to end -- the serialized proving-key
/// queries, the full assignment, (r, s) and the proof bytes -- in hex.
fn prove() {
    use ark_groth16::{create_proof, generate_random_parameters, prepare_verifying_key, verify_proof};
    use zkmember::commitments::pedersen381::{common::*, MerkleTreeCircuit};
    use zkmember::member::Member;
    let mut rng = test_rng();
    let leaf_crh_params = <LeafHash as ark_crypto_primitives::CRH>::setup(&mut rng).unwrap();
    let two_to_one_crh_params = <TwoToOneHash as ark_crypto_primitives::crh::TwoToOneCRH>::setup(&mut rng).unwrap();
    let fixed = chrono::DateTime::parse_from_rfc3339("2024-01-01T00:00:00Z").unwrap().with_timezone(&chrono::Utc);
    let members: Vec<Member> = (1..=3u32).map(|i| {
        let mut m = Member::new(i.to_string().into(), format!("{}@usc.edu", i).into(), None);
        m.join_date = fixed;                                  // the only time-dependent field
        m
    }).collect();
    let tree = new_membership_tree(&leaf_crh_params, &two_to_one_crh_params, &members);
    let index = 1usize;
    let path = tree.generate_proof(index).unwrap();
    let circuit = MerkleTreeCircuit {
        leaf_crh_params: &leaf_crh_params,
        two_to_one_crh_params: &two_to_one_crh_params,
        root: tree.root(),
        leaf_hash: members[index].hash::<LeafHash>(&leaf_crh_params).unwrap(),
        authentication_path: Some(path),
    };
    let pk = generate_random_parameters::<Bls12_381, _, _>(circuit.clone(), &mut rng).unwrap();
    let (r, s) = (Fr::rand(&mut rng), Fr::rand(&mut rng));
    let proof = create_proof(circuit.clone(), &pk, r, s).unwrap();
    let pvk = prepare_verifying_key(&pk.vk);
    let inputs = [tree.root(), circuit.leaf_hash];
    assert!(verify_proof(&pvk, &proof, &inputs).unwrap());
    let mut proof_bytes = Vec::new();
    proof.serialize(&mut proof_bytes).unwrap();
    let mut pk_bytes = Vec::new();
    pk.serialize_uncompressed(&mut pk_bytes).unwrap();
    println!("{{\"generator\":\"rust/parity-driver prove (arkworks 0.3.0)\",\"r\":\"{}\",\"s\":\"{}\",\"proof\":\"{}\",\
              \"proving_key_uncompressed\":\"{}\",\"verified\":true}}",
             fr_limbs(&[r]), fr_limbs(&[s]), hex(&proof_bytes), hex(&pk_bytes));
}
