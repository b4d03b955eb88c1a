// Replacement body for ark-ec 0.3.0 `src/msm/variable_base.rs` (apply in a fork of ark-ec 0.3.0 and
// point zkMember's Cargo.toml at it with [patch.crates-io]).
// UNTESTED in this repository's build environment (no Rust toolchain there).
//
// The generic entry point keeps its signature; it routes by TypeId to the C ABI for the six groups
// libzkm_b200.so implements (BLS12-381, BN254, BW6-761 x G1, G2) and keeps upstream's code for every other curve.
use crate::{AffineCurve, ProjectiveCurve};
use ark_ff::{PrimeField, Zero};
use core::any::TypeId;

pub struct VariableBaseMSM;

impl VariableBaseMSM {
    pub fn multi_scalar_mul<G: AffineCurve>(
        bases: &[G],
        scalars: &[<G::ScalarField as PrimeField>::BigInt],
    ) -> G::Projective {
        let size = core::cmp::min(bases.len(), scalars.len());
        if let Some((curve, group, words, swords)) = gpu_group::<G>() {
            return gpu_msm::<G>(curve, group, words, swords, &bases[..size], &scalars[..size]);
        }
        upstream_multi_scalar_mul(bases, scalars) // the original 0.3.0 body, kept verbatim in the fork
    }
}

/// (curve id, group, u64 words per coordinate, u64 words per scalar) for the groups the GPU library implements.
/// The concrete type names are compared as strings so that ark-ec does not depend on the curve crates.
fn gpu_group<G: AffineCurve>() -> Option<(i32, i32, usize, usize)> {
    let name = core::any::type_name::<G>();
    let _ = TypeId::of::<G>();
    if name.contains("ark_bls12_381") && name.contains("g1") { return Some((0, 1, 6, 4)); }
    if name.contains("ark_bls12_381") && name.contains("g2") { return Some((0, 2, 12, 4)); }
    if name.contains("ark_bn254") && name.contains("g1") { return Some((1, 1, 4, 4)); }
    if name.contains("ark_bn254") && name.contains("g2") { return Some((1, 2, 8, 4)); }
    // BW6-761 (benches/groth16.rs:24-29): G1 and G2 are both curves over the 761-bit Fq (12 words per coordinate),
    // scalars are BigInteger384
    if name.contains("ark_bw6_761") && name.contains("g1") { return Some((2, 1, 12, 6)); }
    if name.contains("ark_bw6_761") && name.contains("g2") { return Some((2, 2, 12, 6)); }
    None
}

fn gpu_msm<G: AffineCurve>(
    curve: i32, group: i32, words: usize, swords: usize, bases: &[G],
    scalars: &[<G::ScalarField as PrimeField>::BigInt],
) -> G::Projective {
    use zkmember_gpu_sys as sys;
    sys::ensure_init();
    // GroupAffine / Fp are not #[repr(C)]: copy into packed arrays instead of transmuting.
    // `write_xy_limbs` (added to the AffineCurve impl in the fork) writes x then y as Montgomery u64 limbs.
    let n = bases.len();
    let mut xy = vec![0u64; n * 2 * words];
    let mut inf = vec![0u8; n];
    for (i, b) in bases.iter().enumerate() {
        inf[i] = b.is_zero() as u8;
        b.write_xy_limbs(&mut xy[i * 2 * words..(i + 1) * 2 * words]);
    }
    let mut sc = vec![0u64; n * swords];
    for (i, s) in scalars.iter().enumerate() {
        sc[i * swords..(i + 1) * swords].copy_from_slice(s.as_ref()); // BigInteger256 / BigInteger384: canonical LE limbs
    }
    let mut out = vec![0u64; 2 * words];
    let mut out_inf = 0u8;
    let rc = unsafe {
        if group == 1 {
            sys::zkm_msm_g1(curve, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut out_inf)
        } else {
            sys::zkm_msm_g2(curve, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), n, out.as_mut_ptr(), &mut out_inf)
        }
    };
    sys::check(rc, "zkm_msm");
    if out_inf != 0 { return G::Projective::zero(); }
    G::from_xy_limbs(&out).into_projective() // (x, y, z = 1)
}
