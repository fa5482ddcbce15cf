"""Mirror of the numeric half of app/utils/shared_pphmm_graphs.py:59-97 and :202-207."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

shared_pphmm_ratio = _ref(_d.shared_pphmm_ratio)
shared_norm_pphmm_ratio = _ref(_d.shared_norm_pphmm_ratio)
pphmm_loc_diffs_pairwise = _ref(_d.pphmm_loc_diffs_pairwise)
