// Additions to a fork of ark-ff 0.3.0 (pin /root/reference/Cargo.lock:229-230).  UNCOMPILED here (no Rust toolchain in
// this repository's build environment); written against the 0.3.0 sources' public items.
//
// Why ark-ff has to be touched at all: the C ABI exchanges field elements as their Montgomery u64 limbs
// (include/zkm_b200.h), and generic code (`G: AffineCurve`, `F: FftField`) cannot reach `Fp384(pub BigInteger384, ..)`'s
// tuple field through any 0.3.0 trait -- `into_repr()` / `write()` give the CANONICAL integer.  Two hidden, defaulted
// methods on `Field` give generic code that access without `transmute` and without assuming a struct layout.
//
// ---------------------------------------------------------------------------------------------------------------
// 1. src/fields/mod.rs -- inside `pub trait Field: ... {`
// ---------------------------------------------------------------------------------------------------------------
//
//     /// zkmember-gpu: append this element's Montgomery limbs (base-prime-field coefficients in order, little-endian
//     /// u64 limbs each) to `out`; returns the number of u64 written.  Default: 0 = "not a supported representation".
//     #[doc(hidden)]
//     fn zkm_write_raw(&self, _out: &mut [u64]) -> usize { 0 }
//     /// zkmember-gpu: rebuild an element from the limbs `zkm_write_raw` produces (no reduction, no conversion).
//     #[doc(hidden)]
//     fn zkm_read_raw(_limbs: &[u64]) -> Option<Self> { None }
//
// ---------------------------------------------------------------------------------------------------------------
// 2. src/fields/macros.rs -- inside `macro_rules! impl_Fp`, in `impl<P: $FpParameters> Field for $Fp<P> {`
//    ($Fp<P> is `pub struct $Fp<P>(pub $BigIntegerType, pub PhantomData<P>)`: `self.0 .0` is `[u64; $limbs]`,
//    the value times R = 2^(64 * $limbs) mod p -- exactly what the C ABI calls "Montgomery limbs")
// ---------------------------------------------------------------------------------------------------------------
//
//     #[inline]
//     fn zkm_write_raw(&self, out: &mut [u64]) -> usize {
//         out[..$limbs].copy_from_slice(&(self.0).0);
//         $limbs
//     }
//     #[inline]
//     fn zkm_read_raw(limbs: &[u64]) -> Option<Self> {
//         if limbs.len() != $limbs { return None; }
//         let mut l = [0u64; $limbs];
//         l.copy_from_slice(limbs);
//         Some($Fp::<P>($BigIntegerType::new(l), PhantomData))     // already Montgomery, already reduced
//     }
//
// ---------------------------------------------------------------------------------------------------------------
// 3. src/fields/models/quadratic_extension.rs -- in `impl<P: QuadExtParameters> Field for QuadExtField<P> {`
//    (G2 of BLS12-381 / BN254: the C ABI orders an Fq2 coordinate c0 then c1)
// ---------------------------------------------------------------------------------------------------------------
//
//     fn zkm_write_raw(&self, out: &mut [u64]) -> usize {
//         let a = self.c0.zkm_write_raw(out);
//         if a == 0 { return 0; }
//         let b = self.c1.zkm_write_raw(&mut out[a..]);
//         if b == 0 { 0 } else { a + b }
//     }
//     fn zkm_read_raw(limbs: &[u64]) -> Option<Self> {
//         let h = limbs.len() / 2;
//         Some(QuadExtField::new(P::BaseField::zkm_read_raw(&limbs[..h])?, P::BaseField::zkm_read_raw(&limbs[h..])?))
//     }
//
// Nothing else in ark-ff changes; every other `Field` implementor keeps the defaults and is never routed to the GPU.

// The block below is real code (not a comment) so that the accessor contract can be unit-tested inside the fork:
// `cargo test -p ark-ff zkm_raw` after pasting sections 1-3.
#[cfg(test)]
mod zkm_raw_tests {
    use crate::{Field, One, PrimeField, UniformRand};
    fn roundtrip<F: Field>() {
        let mut rng = ark_std::test_rng();
        let mut buf = [0u64; 32];
        for _ in 0..100 {
            let x = F::rand(&mut rng);
            let n = x.zkm_write_raw(&mut buf);
            assert!(n > 0);
            assert_eq!(F::zkm_read_raw(&buf[..n]).unwrap(), x);
        }
    }
    fn one_is_r<F: PrimeField>() {
        // the raw limbs of one() are R mod p: the definition of the Montgomery form the C ABI documents
        let mut buf = [0u64; 16];
        let n = F::one().zkm_write_raw(&mut buf);
        assert_eq!(&buf[..n], <F::Params as crate::FpParameters>::R.as_ref());
    }
    #[test]
    fn zkm_raw() {
        roundtrip::<crate::test_field::Fq>();            // any Fp the crate's own tests instantiate
        one_is_r::<crate::test_field::Fq>();
    }
}
