"""Python mirror of package g2pubs (public keys in G2, signatures in G1; g2pubs/bls.go:18-278) on top of the engine;
BASELINE.json's first configuration (g2pubs Sign / Verify) is this module.  See g1pubs.py / _pubs.py."""
from . import _pubs

_pubs.build(globals(), 2, "g2pubs")
