"""B200-native MPBP message-update engine behind the MatrixProductBP.jl API surface.
Import through the alias package ``mpbp_b200`` (see /mpbp_b200/__init__.py)."""
