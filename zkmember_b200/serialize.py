"""arkworks `CanonicalSerialize` (compressed) for affine points and Groth16 proofs -- the wire format zkMember
ships proofs in (/root/reference/src/main.rs:164-169,217).

Restated from ark-ec 0.3.0 `GroupAffine::serialize` (src/models/short_weierstrass_jacobian.rs) and ark-ff /
ark-serialize 0.3.0 `Fp::serialize_with_flags`, `QuadExtField::serialize_with_flags`, `SWFlags`:

  * a point is its x coordinate, canonical (`into_repr`) little-endian, in ceil((MODULUS_BITS + 2) / 8) bytes;
  * the two flag bits live in the top bits of the LAST byte: bit 7 = "y is the larger of {y, -y}"
    (`SWFlags::PositiveY`), bit 6 = point at infinity (then x = 0);
  * an Fp2 coordinate writes c0 without flags (ceil(MODULUS_BITS / 8) bytes) and c1 with them; Fp2 elements are
    ordered by c1 first, then c0;
  * `Proof { a, b, c }` is a || b || c  (BLS12-381: 48 + 96 + 48 = 192 bytes; BW6-761: 96 + 96 + 96).

Host-side glue on a handful of field elements per proof -- nothing here is on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from .msm import AffinePoint, L64

# base-field moduli (published curve parameters; Montgomery radix R = 2^(64 * limbs))
FQ_MODULUS = {
    _lib.CURVE_BLS12_381: 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    _lib.CURVE_BN254: 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    _lib.CURVE_BW6_761: 0x122e824fb83ce0ad187c94004faff3eb926186a81d14688528275ef8087be41707ba638e584e91903cebaff25b423048689c8ed12f9fd9071dcd3dc73ebff2e98a116c25667a8f8160cf8aeeaf0a437e6913e6870000082f49d00000000008b,
}
FR_MODULUS = {
    _lib.CURVE_BLS12_381: 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    _lib.CURVE_BN254: 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    _lib.CURVE_BW6_761: 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
}


def _fq_from_mont(curve: int, limbs: np.ndarray) -> int:
    p = FQ_MODULUS[curve]
    v = int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little")
    return v * pow(1 << (64 * L64[curve]), -1, p) % p


def _coords(point: AffinePoint):
    """(x, y) as tuples of canonical Fq integers (one entry for Fq, two -- c0, c1 -- for Fq2)."""
    l = L64[point.curve]
    words = len(point.xy) // 2
    deg = words // l
    x = tuple(_fq_from_mont(point.curve, point.xy[i * l:(i + 1) * l]) for i in range(deg))
    y = tuple(_fq_from_mont(point.curve, point.xy[words + i * l:words + (i + 1) * l]) for i in range(deg))
    return x, y


def serialize_affine(point: AffinePoint) -> bytes:
    """`GroupAffine::serialize` (compressed)."""
    p = FQ_MODULUS[point.curve]
    bits = p.bit_length()
    plain, flagged = (bits + 7) // 8, (bits + 2 + 7) // 8
    x, y = _coords(point)
    if point.infinity:
        x = tuple(0 for _ in x)
        flag = 1 << 6
    else:
        neg = tuple((p - c) % p for c in y)
        # Fp: larger canonical integer; Fp2: compare c1 first, then c0
        flag = (1 << 7) if tuple(reversed(y)) > tuple(reversed(neg)) else 0
    out = bytearray()
    for c in x[:-1]:
        out += c.to_bytes(plain, "little")
    last = bytearray(x[-1].to_bytes(flagged, "little"))
    last[-1] |= flag
    return bytes(out + last)


@dataclass
class Proof:
    """ark_groth16::Proof { a: G1Affine, b: G2Affine, c: G1Affine }."""
    a: AffinePoint
    b: AffinePoint
    c: AffinePoint

    def serialize(self) -> bytes:
        return serialize_affine(self.a) + serialize_affine(self.b) + serialize_affine(self.c)
