"""Mirror of app/utils/similarity_matrix_constructor.py (reference import path)."""
from ...dropin import SimilarityMat_Constructor  # noqa: F401
