"""General recommenders of the MF family on the sm_100a hot path: MF (BPRMF / pointwise),
MLP, NeuMF, LightGCN, NGCF.  Resolved by name from main.py like the reference (main.py:30-40)."""
