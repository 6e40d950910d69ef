"""lah_b200 — Blackwell-native Decentralized Mixture-of-Experts engine (see README.md / DESIGN.md)."""
__version__ = "0.1.0"
