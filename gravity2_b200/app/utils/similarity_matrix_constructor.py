"""Mirror of app/utils/similarity_matrix_constructor.py (reference import path)."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

SimilarityMat_Constructor = _ref(_d.SimilarityMat_Constructor)
