"""Mirror of app/utils/dist_mat_to_tree.py (Mayne941/gravity2): same name, same signature, computed on a B200."""
from gravity2_b200.dropin import DistMat2Tree  # noqa: F401
