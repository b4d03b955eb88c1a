// Replacement for the four in-place transforms of ark-poly 0.3.0 `src/domain/radix2/mod.rs`
// (`impl<F: FftField> EvaluationDomain<F> for Radix2EvaluationDomain<F>`), to be applied in a fork.
// UNTESTED in this repository's build environment (no Rust toolchain there).
//
//   fn fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>)
//   fn ifft_in_place<T: DomainCoeff<F>>(&self, evals: &mut Vec<T>)
//   fn coset_fft_in_place / coset_ifft_in_place        (trait defaults overridden)
//
// Only T == F == Fr of BLS12-381 / BN254 / BW6-761 is routed to the GPU; anything else keeps upstream's code.
// Returns (curve id, u64 limbs per element).
fn gpu_field<F: 'static>() -> Option<(i32, usize)> {
    let name = core::any::type_name::<F>();
    if name.contains("ark_bls12_381") && name.contains("Fr") { return Some((0, 4)); }
    if name.contains("ark_bn254") && name.contains("Fr") { return Some((1, 4)); }
    // ark_bw6_761::Fr is a re-export of ark_bls12_377::Fq (FqParameters): match the concrete parameter type
    if name.contains("ark_bls12_377") && name.contains("FqParameters") { return Some((2, 6)); }
    None
}

fn gpu_ntt<F: FftField>(dom: &Radix2EvaluationDomain<F>, v: &mut Vec<F>, inverse: bool, coset: bool, curve: i32, l: usize) {
    use zkmember_gpu_sys as sys;
    sys::ensure_init();
    v.resize(dom.size(), F::zero());                       // upstream: coeffs.resize(self.size(), T::zero())
    // Fp256 / Fp384 are not #[repr(C)]: stage through a packed u64 buffer (l limbs per element, Montgomery).
    let mut buf = vec![0u64; l * v.len()];
    for (i, e) in v.iter().enumerate() { buf[l * i..l * (i + 1)].copy_from_slice(e.montgomery_limbs()); }
    let rc = unsafe { sys::zkm_ntt(curve, buf.as_mut_ptr(), dom.log_size_of_group, inverse as i32, coset as i32) };
    sys::check(rc, "zkm_ntt");
    for (i, e) in v.iter_mut().enumerate() { *e = F::from_montgomery_limbs(&buf[l * i..l * (i + 1)]); }
}
// in the impl block:
//   fn fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>) {
//       if TypeId::of::<T>() == TypeId::of::<F>() { if let Some((c, l)) = gpu_field::<F>() {
//           return gpu_ntt(self, cast_vec_mut::<T, F>(coeffs), false, false, c, l); } }
//       /* upstream body */
//   }
//   … same for ifft_in_place (true,false), coset_fft_in_place (false,true), coset_ifft_in_place (true,true).
