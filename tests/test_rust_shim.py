"""The Rust side cannot be compiled here (no cargo/rustc); what CAN be pinned on CPU: the modulus constants the
patches route by equal the curve parameters, and every `extern "C"` item of the -sys crate is declared in
include/zkm_b200.h with the same number of parameters."""
import os
import re

from oracle.py.params import BLS12_381, BN254, BW6_761

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"BLS12_381_FQ": BLS12_381.fq, "BLS12_381_FR": BLS12_381.fr, "BN254_FQ": BN254.fq, "BN254_FR": BN254.fr,
        "BW6_761_FQ": BW6_761.fq, "BW6_761_FR": BW6_761.fr}


def test_routing_moduli_in_the_patches_equal_the_curve_parameters():
    seen = 0
    for f in ("rust/patches/ark_ec_variable_base.rs", "rust/patches/ark_poly_radix2.rs"):
        s = open(os.path.join(ROOT, f)).read()
        for name, body in re.findall(r"const (\w+): \[u64; \d+\] = \[(.*?)\];", s, re.S):
            limbs = [int(x, 16) for x in re.findall(r"0x[0-9a-f]+", body)]
            assert sum(l << (64 * i) for i, l in enumerate(limbs)) == WANT[name].modulus, (f, name)
            assert len(limbs) == WANT[name].limbs64
            seen += 1
    assert seen == 9


def test_sys_crate_declares_only_functions_of_the_header_with_matching_arity():
    header = open(os.path.join(ROOT, "include", "zkm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(zkm_\w+)\s*\(([^;]*?)\)\s*;", header, re.S)}
    lib = open(os.path.join(ROOT, "rust", "zkmember-gpu-sys", "src", "lib.rs")).read()
    block = lib[lib.index('extern "C" {'):]
    block = block[:block.index("\n}")]
    fns = re.findall(r"pub fn (zkm_\w+)\s*\((.*?)\)\s*(?:->\s*[\w*: ]+)?;", block, re.S)
    assert len(fns) >= 20
    arity = lambda params: 0 if params.strip() in ("", "void") else len([p for p in params.split(",") if p.strip()])
    for name, params in fns:
        assert name in decl, name
        assert arity(params) == arity(decl[name]), name
