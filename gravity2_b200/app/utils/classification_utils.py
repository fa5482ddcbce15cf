"""Mirror of the numeric half of app/utils/classification_utils.py:10-68."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

PairwiseSimilarityScore_Cutoff_Dict_Constructor = _ref(_d.PairwiseSimilarityScore_Cutoff_Dict_Constructor)
max_similarity = _ref(_d.max_similarity)
