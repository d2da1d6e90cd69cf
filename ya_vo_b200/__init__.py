"""ya_vo_b200 — B200-native (sm_100a) FAST + BRIEF + Hamming front end behind YA_VO's
FastDetector / Brief / Image interface.  See DESIGN.md."""
__version__ = "0.1.0"
