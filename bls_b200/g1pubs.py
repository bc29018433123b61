"""Python mirror of package g1pubs (public keys in G1, signatures in G2; g1pubs/bls.go): same names, argument meaning
and results as the reference -- Sign, SignWithDomain, PrivToPub, Verify, VerifyWithDomain, AggregateSignatures,
AggregatePublicKeys, (*Signature).VerifyAggregate / VerifyAggregateCommon / ...WithDomain, Serialize / Deserialize*,
DeriveSecretKey, RandKey (g1pubs/bls.go:18-311) -- with every curve operation executed by libb381.so through the C ABI
(bls_b200/_pubs.py).  VerifyBatch / VerifyWithDomainBatch are additions over wire-format inputs."""
from . import _pubs

_pubs.build(globals(), 1, "g1pubs")
