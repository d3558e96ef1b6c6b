"""Social recommenders on the fused triplet path."""
