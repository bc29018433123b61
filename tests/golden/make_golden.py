#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into tests/golden/ref_kats.json.

Run in the build container only (reads /root/reference/*_test.go, which does not exist on the
GPU box); the JSON is committed.  Only numbers are extracted - the checks that use them are
written in tests/test_oracle_golden.py, each citing the reference test it mirrors.
"""
import json, pathlib, re

REF = pathlib.Path("/root/reference")
OUT = pathlib.Path(__file__).with_name("ref_kats.json")


def func_body(src, name):
    m = re.search(r"func %s\(t \*testing\.T\) \{\n(.*?)\n\}\n" % name, src, re.S)
    assert m, name
    return m.group(1)


def strings(body):
    """All FQReprFromString("...", base) literals in order -> canonical ints (hex strings)."""
    out = []
    for s, base in re.findall(r'FQReprFromString\("([0-9a-fA-F]+)", (\d+)\)', body):
        out.append(hex(int(s, int(base))))
    return out


def reprs(body):
    """All FQRepr{...} literals (6 x u64, least significant first) -> ints (hex strings)."""
    out = []
    for lit in re.findall(r"FQRepr\{([^}]*)\}", body):
        limbs = [int(x, 0) for x in lit.replace("\n", " ").split(",") if x.strip()]
        assert len(limbs) == 6
        out.append(hex(sum(l << (64 * i) for i, l in enumerate(limbs))))
    return out


def struct_cases(body, nfields):
    nums = [int(x) for x in re.findall(r"(?:\w+:\s*)?(\d+),", body)]
    assert len(nums) % nfields == 0
    return [nums[i:i + nfields] for i in range(0, len(nums), nfields)]


k = {}
src = (REF / "pairing_test.go").read_text()
k["pairing_g1_g2"] = {"cite": "pairing_test.go:9-58", "coeffs": strings(src)[:12]}

src = (REF / "fqrepr_test.go").read_text()
hi, lo, exp = reprs(func_body(src, "TestMontReduce"))
k["mont_reduce"] = {"cite": "fqrepr_test.go:136-147", "hi": hi, "lo": lo, "expected": exp}

src = (REF / "fq_test.go").read_text()
k["fq_inverse_input"] = {"cite": "fq_test.go:189-207", "value": strings(func_body(src, "TestInverse"))[0]}

src = (REF / "fq2_test.go").read_text()
k["fq2"] = {"cite": "fq2_test.go:71-246"}
for name in ["TestFQ2Squaring", "TestFQ2Mul", "TestFQ2Inverse", "TestFQ2Addition", "TestFQ2Subtraction",
             "TestFQ2Negation", "TestFQ2Doubling", "TestFQ2FrobeniusMap", "TestFQ2Sqrt"]:
    k["fq2"][name] = strings(func_body(src, name))

src = (REF / "g1_test.go").read_text()
k["g1_double"] = {"cite": "g1_test.go:62-79", "values": reprs(func_body(src, "TestG1DoublingCorrectness"))}
k["g1_add"] = {"cite": "g1_test.go:81-104", "values": reprs(func_body(src, "TestG1AdditionCorrectness"))}

src = (REF / "primitivefuncs_test.go").read_text()
body = func_body(src, "TestSubWithCarry")
k["sub_with_borrow"] = {"cite": "primitivefuncs_test.go:25-101",
                        "cases": struct_cases(body[:body.index("for _, c")], 5)}
body = func_body(src, "TestAddWithCarry")
k["add_with_carry"] = {"cite": "primitivefuncs_test.go:110-186",
                       "cases": struct_cases(body[:body.index("for _, c")], 5)}
body = func_body(src, "TestMACWithCarry")
k["mac_with_carry"] = {"cite": "primitivefuncs_test.go:188-240",
                       "cases": struct_cases(body[:body.index("for _, c")], 6)}
body = func_body(src, "TestMultiplyFQReprOverflow")
f0, f1 = reprs(body)
lo_, hi_ = re.findall(r'SetString\("(\d+)", 10\)', body)
k["multiply_fq_repr"] = {"cite": "primitivefuncs_test.go:244-262", "f0": f0, "f1": f1,
                         "lo": hex(int(lo_)), "hi": hex(int(hi_))}

# constants that pin the field (fq.go:26,29; stub_fallback.go:59; g1.go:25-29; g2.go:26-29)
fq = (REF / "fq.go").read_text()
g1 = (REF / "g1.go").read_text()
g2 = (REF / "g2.go").read_text()
k["constants"] = {
    "cite": "fq.go:26,29; g1.go:25-29; g2.go:26-29",
    "q": hex(int(re.search(r'QFieldModulus, _ = FQReprFromString\("(\d+)"', fq).group(1))),
    "r2": hex(int(re.search(r'FQR2, _ = FQReprFromString\("(\d+)"', fq).group(1))),
    "g1_gen": [hex(int(x)) for x in re.findall(r'g1Generator[XY], _ = FQReprFromString\("(\d+)", 10\)', g1)],
    "b_coeff_mont": reprs(re.search(r"var BCoeff = .*", g1).group(0))[0],
    "g2_gen_xc1_xc0_yc1_yc0": [hex(int(x, 16)) for x in
                               re.findall(r'g2Generator\w+, _ = FQReprFromString\("([0-9a-f]+)", 16\)', g2)],
}
# Frobenius tables (Montgomery-form literals): fq2.go:149-152, fq6.go:144-208, fq12.go:122-168
def table(src, name):
    m = re.search(r"var %s = \[\d+\]FQ2?\{(.*?)\n\}\n" % name, src, re.S)
    return m.group(1)
fq6 = (REF / "fq6.go").read_text()
fq12 = (REF / "fq12.go").read_text()
fq2 = (REF / "fq2.go").read_text()
k["frobenius_mont"] = {
    "cite": "fq2.go:149-152; fq6.go:144-208; fq12.go:122-168",
    "fq2_c1_1": reprs(table(fq2, "frobeniusCoeffFQ2c1"))[0],
    "fq6_c1": reprs(table(fq6, "frobeniusCoeffFQ6c1")),
    "fq6_c2": reprs(table(fq6, "frobeniusCoeffFQ6c2")),
    "fq12_c1_from_1": reprs(table(fq12, "frobeniusCoeffFQ12c1")),
}
# hash-to-curve, key derivation and serialisation vectors (hash_test.go:12-82, g1pubs/bls_test.go:409-433, g2pubs/bls_test.go:336-347)
ht = (REF / "hash_test.go").read_text()
hx = lambda name: hex(int(re.search(r'%s, _ = bls.FQReprFromString\("([0-9a-f]+)", 16\)' % name, ht).group(1), 16))
g1t = (REF / "g1pubs" / "bls_test.go").read_text()
g2t = (REF / "g2pubs" / "bls_test.go").read_text()
k["hash"] = {
    "cite": "hash_test.go:12-82; g1pubs/bls_test.go:409-433; g2pubs/bls_test.go:336-347",
    "message": "the message to be signed",
    "hash_g1": [hx("expectedG1X"), hx("expectedG1Y")],
    "hash_g2_xc0_xc1_yc0_yc1": [hx("expectedG2c0X"), hx("expectedG2c1X"), hx("expectedG2c0Y"), hx("expectedG2c1Y")],
    "hash_g2_with_domain_zero_compressed": re.search(r'expectedSerializedG2, _ = hex.DecodeString\("([0-9a-f]+)"\)', ht).group(1),
    "derive_secret_key_in": re.search(r'copy\(secKeyIn\[:\], \[\]byte\("(\d+)"\)\)', g1t).group(1),
    "derive_secret_key_out": hex(int(re.search(r'FRReprFromString\("([0-9a-f]+)", 16\)', g1t).group(1), 16)),
    "invalid_g1_pubkey": re.search(r'func TestPubkeyDeserializeInvalid.*?unexpectedPub := "([0-9a-f]+)"', g1t, re.S).group(1),
    "invalid_g2_pubkey": re.search(r'func TestPubkeyDeserializeInvalid.*?unexpectedPub := "([0-9a-f]+)"', g2t, re.S).group(1),
}
OUT.write_text(json.dumps(k, indent=1) + "\n")
print("wrote", OUT, {a: (len(b) if hasattr(b, "__len__") else b) for a, b in k.items()})
