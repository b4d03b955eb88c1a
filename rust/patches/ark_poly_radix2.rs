// Replacement for the four in-place transforms of ark-poly 0.3.0 `src/domain/radix2/mod.rs`
// (`impl<F: FftField> EvaluationDomain<F> for Radix2EvaluationDomain<F>`; pin /root/reference/Cargo.lock:338-339), to
// be applied in a fork selected with [patch.crates-io].  UNCOMPILED here (no Rust toolchain in this repository's build
// environment); complete source against the 0.3.0 public API + the raw-limb accessors of ark_ff_raw_limbs.rs.
//
// Only `T == F == Fr` of BLS12-381 / BN254 / BW6-761 goes to the GPU (Groth16's witness_map and Marlin's AHP transform
// vectors of field elements); `T = G::Projective` (DomainCoeff for group elements) and every other field keep
// upstream's code.  Cargo.toml of the fork:  zkmember-gpu-sys = { path = "<repo>/rust/zkmember-gpu-sys" }
use ark_ff::{FftField, Field, FpParameters, PrimeField};
use core::any::TypeId;
use zkmember_gpu_sys as sys;

const BLS12_381_FR: [u64; 4] = [0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48];
const BN254_FR: [u64; 4] = [0x43e1f593f0000001, 0x2833e84879b97091, 0xb85045b68181585d, 0x30644e72e131a029];
const BW6_761_FR: [u64; 6] = [0x8508c00000000001, 0x170b5d4430000000, 0x1ef3622fba094800, 0x1a22d9f300f5138f, 0xc63b05c06ca1493b, 0x01ae3a4617c510ea];

/// (curve id, u64 limbs per element) when F is one of the three scalar fields, by its modulus (exact, no type names).
fn gpu_field<F: FftField>() -> Option<(i32, usize)> {
    if F::extension_degree() != 1 { return None; }
    let m = <<F::BasePrimeField as PrimeField>::Params as FpParameters>::MODULUS;
    let m = m.as_ref();
    if m == &BLS12_381_FR[..] { return Some((sys::ZKM_CURVE_BLS12_381, 4)); }
    if m == &BN254_FR[..] { return Some((sys::ZKM_CURVE_BN254, 4)); }
    if m == &BW6_761_FR[..] { return Some((sys::ZKM_CURVE_BW6_761, 6)); }   // ark_bw6_761::Fr = ark_bls12_377::Fq
    None
}

/// `&mut Vec<T>` as `&mut Vec<F>` when the two types are the same type (checked with TypeId: no layout assumption).
fn same_type_mut<T: 'static, F: 'static>(v: &mut Vec<T>) -> Option<&mut Vec<F>> {
    if TypeId::of::<T>() == TypeId::of::<F>() {
        // SAFETY: T and F are the same type, so this is the identity cast.
        Some(unsafe { &mut *(v as *mut Vec<T> as *mut Vec<F>) })
    } else {
        None
    }
}

fn gpu_ntt<F: FftField>(size: usize, log_size: u32, v: &mut Vec<F>, inverse: bool, coset: bool, curve: i32, limbs: usize) {
    sys::ensure_init();
    v.resize(size, F::zero());                             // upstream: coeffs.resize(self.size(), T::zero())
    // Fp256 / Fp384 are not #[repr(C)]: stage through a packed u64 buffer (Montgomery limbs)
    let mut buf = vec![0u64; limbs * v.len()];
    for (i, e) in v.iter().enumerate() {
        let n = e.zkm_write_raw(&mut buf[limbs * i..limbs * (i + 1)]);
        assert_eq!(n, limbs, "zkmember-gpu: unexpected element width");
    }
    sys::check(unsafe { sys::zkm_ntt(curve, buf.as_mut_ptr(), log_size, inverse as i32, coset as i32) }, "zkm_ntt");
    for (i, e) in v.iter_mut().enumerate() {
        *e = F::zkm_read_raw(&buf[limbs * i..limbs * (i + 1)]).expect("zkmember-gpu: field without raw-limb access");
    }
}

// In `impl<F: FftField> EvaluationDomain<F> for Radix2EvaluationDomain<F>` the four methods become (the upstream
// bodies move, unchanged, into `*_upstream` private methods of Radix2EvaluationDomain<F>):
//
//     fn fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>) {
//         if let (Some((c, l)), Some(v)) = (gpu_field::<F>(), same_type_mut::<T, F>(coeffs)) {
//             return gpu_ntt(self.size(), self.log_size_of_group, v, false, false, c, l);
//         }
//         self.fft_in_place_upstream(coeffs)
//     }
//     fn ifft_in_place<T: DomainCoeff<F>>(&self, evals: &mut Vec<T>) {
//         if let (Some((c, l)), Some(v)) = (gpu_field::<F>(), same_type_mut::<T, F>(evals)) {
//             return gpu_ntt(self.size(), self.log_size_of_group, v, true, false, c, l);
//         }
//         self.ifft_in_place_upstream(evals)
//     }
//     fn coset_fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>) {      // trait default overridden
//         if let (Some((c, l)), Some(v)) = (gpu_field::<F>(), same_type_mut::<T, F>(coeffs)) {
//             return gpu_ntt(self.size(), self.log_size_of_group, v, false, true, c, l);
//         }
//         Self::distribute_powers(coeffs, F::multiplicative_generator());
//         self.fft_in_place_upstream(coeffs)
//     }
//     fn coset_ifft_in_place<T: DomainCoeff<F>>(&self, evals: &mut Vec<T>) {      // trait default overridden
//         if let (Some((c, l)), Some(v)) = (gpu_field::<F>(), same_type_mut::<T, F>(evals)) {
//             return gpu_ntt(self.size(), self.log_size_of_group, v, true, true, c, l);
//         }
//         self.ifft_in_place_upstream(evals);
//         Self::distribute_powers(evals, self.generator_inv)
//     }
//
// `DomainCoeff<F>` requires `'static` in neither crate version, so the fork adds `T: 'static` to the four method
// signatures' where-clauses (every implementor in arkworks -- F itself and the projective groups -- is 'static).
