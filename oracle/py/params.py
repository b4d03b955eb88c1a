"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Curve / field parameters for the two curve families the hot path covers
(BLS12-381 and BN254), restated from the published parameters of
ark-bls12-381 0.3.0 / ark-bn254 0.3.0 (pins: /root/reference/Cargo.lock:124-125,
135-136).  Everything here is re-derived and checked
numerically by tests/test_oracle_constants.py (primality, generator order,
on-curve, two-adic root order, Montgomery constants against the published limb
values of the upstream crates).

Parity status: PARITY UNPINNED against the arkworks *binaries* (no Rust
toolchain, no vendored crate source, no golden vectors in the reference --
SURVEY.md section 8c).  What is pinned: the mathematics (MSM results are unique
affine group elements, NTT outputs are unique field elements) and the Montgomery
representation (R = 2^(64*limbs), published limb constants).
"""
from dataclasses import dataclass, field


@dataclass(frozen=True)
class FieldParams:
    name: str
    modulus: int
    limbs64: int                 # ark-ff BigInteger{256,384} width in u64 limbs
    generator: int = 0           # F::multiplicative_generator() (FftParameters::GENERATOR)
    two_adicity: int = 0
    two_adic_root: int = 0       # generator^((p-1)/2^two_adicity)

    @property
    def bits(self) -> int:
        return self.modulus.bit_length()

    @property
    def R(self) -> int:          # Montgomery radix, ark-ff: 2^(64*limbs)
        return 1 << (64 * self.limbs64)

    @property
    def limbs32(self) -> int:
        return 2 * self.limbs64

    @property
    def inv64(self) -> int:      # -p^{-1} mod 2^64   (ark-ff FpParameters::INV)
        return (-pow(self.modulus, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        return (-pow(self.modulus, -1, 1 << 32)) % (1 << 32)

    def to_mont(self, x: int) -> int:
        return (x * self.R) % self.modulus

    def from_mont(self, x: int) -> int:
        return (x * pow(self.R, -1, self.modulus)) % self.modulus


@dataclass(frozen=True)
class CurveParams:
    name: str
    curve_id: int                # matches include/zkm_b200.h ZKM_CURVE_*
    fq: FieldParams
    fr: FieldParams
    b_g1: int                    # y^2 = x^3 + b
    g1: tuple                    # generator (x, y)
    b_g2: object                 # (c0, c1) in Fq2 = Fq[u]/(u^2+1); an int when G2 lives over Fq (BW6-761)
    g2: tuple                    # ((x.c0, x.c1), (y.c0, y.c1)); (x, y) when G2 lives over Fq
    g2_over_fq: bool = False     # BW6-761: G2 is a second curve over the SAME prime field (M-twist, degree 1)

    def coord_degree(self, g: int) -> int:
        """Fq elements per coordinate of group g (1 | 2)."""
        return 2 if (g == 2 and not self.g2_over_fq) else 1


# ----------------------------------------------------------------------------- BLS12-381
BLS12_381_FQ = FieldParams(
    name="bls12_381_fq",
    modulus=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    limbs64=6,
)
BLS12_381_FR = FieldParams(
    name="bls12_381_fr",
    modulus=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    limbs64=4,
    generator=7,
    two_adicity=32,
    two_adic_root=0x16a2a19edfe81f20d09b681922c813b4b63683508c2280b93829971f439f0d2b,
)
BLS12_381 = CurveParams(
    name="bls12_381",
    curve_id=0,
    fq=BLS12_381_FQ,
    fr=BLS12_381_FR,
    b_g1=4,
    g1=(
        0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
    ),
    b_g2=(4, 4),
    g2=(
        (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
         0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
         0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be),
    ),
)

# ----------------------------------------------------------------------------- BN254
BN254_FQ = FieldParams(
    name="bn254_fq",
    modulus=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    limbs64=4,
)
BN254_FR = FieldParams(
    name="bn254_fr",
    modulus=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    limbs64=4,
    generator=5,
    two_adicity=28,
    two_adic_root=0x2a3c09f0a58a7e8500e0a7eb8ef62abc402d111e41112ed49bd61b6e725b19f0,
)
_BN254_Q = BN254_FQ.modulus


def _fq2_inv(a, q):
    c0, c1 = a
    n = pow((c0 * c0 + c1 * c1) % q, -1, q)
    return (c0 * n % q, (-c1 * n) % q)


_bn_b2 = _fq2_inv((9, 1), _BN254_Q)
BN254 = CurveParams(
    name="bn254",
    curve_id=1,
    fq=BN254_FQ,
    fr=BN254_FR,
    b_g1=3,
    g1=(1, 2),
    b_g2=(3 * _bn_b2[0] % _BN254_Q, 3 * _bn_b2[1] % _BN254_Q),   # 3/(9+u)
    g2=(
        (10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531),
    ),
)

# ----------------------------------------------------------------------------- BW6-761
# The other pairing curve zkMember instantiates (/root/reference/benches/groth16.rs:24-29,
# /root/reference/benches/marlin.rs:40-73; ark-bw6-761 0.3.0).  Fr is the base field of BLS12-377
# (ark-bw6-761 re-exports ark_bls12_377::Fq as Fr), Fq is the 761-bit prime of the BW6 family
# q = (t^2 + 3 y^2) / 4 with t = x^5 - 3x^4 + 3x^3 - x + 3 + 13 r, y = (x^5 - 3x^4 + 3x^3 - x + 3)/3 + 9 r,
# x = 0x8508c00000000001 (checked by tests/test_oracle_constants.py).  G1: y^2 = x^3 - 1, G2: y^2 = x^3 + 4,
# BOTH over Fq.  FFT constants of Fr as published in ark-bls12-377 0.3.0 src/fields/fq.rs:
# GENERATOR = -5, TWO_ADICITY = 46, TWO_ADIC_ROOT_OF_UNITY = (-5)^((r-1)/2^46); their Montgomery limbs
# (0xfc0b8000000002fa, ... / 0x1c104955744e6e0f, ...) are re-derived in the tests.
# The group generators below are NOT ark-bw6-761's G1/G2_GENERATOR constants (not needed by the hot path:
# an MSM never touches the generator or the coefficient b); they are the cofactor-cleared points over the
# smallest x >= 2 with a curve point (smaller y), used only to synthesise test / bench bases.
BW6_761_FQ = FieldParams(
    name="bw6_761_fq",
    modulus=0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b,
    limbs64=12,
)
BW6_761_FR = FieldParams(
    name="bw6_761_fr",
    modulus=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
    limbs64=6,
    generator=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001 - 5,
    two_adicity=46,
    two_adic_root=pow(0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001 - 5,
                      (0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001 - 1) >> 46,
                      0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001),
)
BW6_761 = CurveParams(
    name="bw6_761",
    curve_id=2,
    fq=BW6_761_FQ,
    fr=BW6_761_FR,
    b_g1=BW6_761_FQ.modulus - 1,
    g1=(
        0xd82cbf66753123ed25942ffadbec116b901330673728468b1653febae12aa13a5d68dc240a36cfbe185365abc6cb0cc5042c14be9179f0c6c05fc952c93a806d5316c2b601db66bd557011eb2c7dd0c1891418e3ce0e512da946c2ca98c56f,
        0xa62fd67fdd91e327a96c02bc80385547a171b11241a2653b54d7359cd7569806b159fd05975390f644cd4d4d121918f1f84be0e364c557f196bd4095e732d987ca22009ba7577b80aaa35b641488679ed9ef0d43b32e776ad507137f20a2dd,
    ),
    b_g2=4,
    g2=(
        0x4cd0ed4bbb4bad28e9646093e4c6ab32a3a80f35265437deef8f50aa1221f459b249d724b2c155e2bf40a492ead210323e3f1c3e6991b9bcabe9da05882daf12d84f49c477fd322fc532b59d18f40b4cc45de6fbd67847acac591e8c5a93fa,
        0x5bbdc19380f7c707f6fe8680ef10cd46fa210a92bc4f56f1b92ab610ecb3fd508160dc51bab3ee5072aa3dedbe0766414556817439a0fbc33df16a239fc281edec5df53182dcf168c9171615a79353ac90858b4a2f2d20aeae7ededaead5b0,
    ),
    g2_over_fq=True,
)

CURVES = {c.name: c for c in (BLS12_381, BN254, BW6_761)}
CURVES_BY_ID = {c.curve_id: c for c in (BLS12_381, BN254, BW6_761)}
