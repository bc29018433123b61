#!/usr/bin/env python3
"""Generate bls_b200/hash_params.json: the numeric parameters of the hash-to-curve suites the reference
implements (isogeny map coefficients, SWU curve constants, psi constants, cofactor).  They are numeric
facts of the IETF draft cipher suite; reference: hash.go:115-301, hash.go:333-339, g1.go:614-618,
g2.go:883-918, g2.go:133.  Run in the build container only (reads /root/reference); the JSON is committed."""
import json, pathlib, re
REF = pathlib.Path("/root/reference")
hash_go, g1_go, g2_go = (REF / "hash.go").read_text(), (REF / "g1.go").read_text(), (REF / "g2.go").read_text()


def fq_list(name, src):
    m = re.search(r"var %s = \[\]FQ\{(.*?)\n\}" % name, src, re.S)
    return re.findall(r'fqReprFromHexUnchecked\("([0-9a-f]+)"\)', m.group(1))


def fq2_list(name, src):
    h = fq_list(name.replace("FQ", "FQ2"), src.replace("[]FQ2{", "[]FQ{")) if False else None
    m = re.search(r"var %s = \[\]FQ2\{(.*?)\n\}" % name, src, re.S)
    h = re.findall(r'fqReprFromHexUnchecked\("([0-9a-f]+)"\)', m.group(1))
    return [list(p) for p in zip(h[0::2], h[1::2])]


def single(name, src):
    return re.search(r'var %s = FQReprToFQ\(fqReprFromHexUnchecked\("([0-9a-f]+)"\)\)' % name, src).group(1)


out = {"iso11": {n: fq_list(n, hash_go) for n in ("xNum11", "xDen11", "yNum11", "yDen11")},
       "iso3": {n: fq2_list(n, hash_go) for n in ("xNum3", "xDen3", "yNum3", "yDen3")}}
m = re.search(r"var iwsc = NewFQ2\((.*?)\n\)", hash_go, re.S)
out["iwsc"] = re.findall(r'"([0-9a-f]+)"', m.group(1))
out["kQiX"], out["kQiY"] = single("kQiX", hash_go), single("kQiY", hash_go)
out["ellPA"] = re.search(r'var ellPARepr, _ = FQReprFromString\("([0-9a-f]+)", 16\)', g1_go).group(1)
out["ellPB"] = re.search(r'var ellPBRepr, _ = FQReprFromString\("([0-9a-f]+)", 16\)', g1_go).group(1)
out["g2_cofactor"] = re.search(r'var g2Cofactor, _ = new\(big.Int\).SetString\("([0-9a-f]+)", 16\)', g2_go).group(1)
out["ell2pA"], out["ell2pB"] = [0, 240], [1012, 1012]           # g2.go:883-891
p = pathlib.Path(__file__).resolve().parent.parent / "bls_b200" / "hash_params.json"
p.write_text(json.dumps(out, indent=1) + "\n")
print("wrote", p)
