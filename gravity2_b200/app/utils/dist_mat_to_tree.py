"""Mirror of app/utils/dist_mat_to_tree.py (Mayne941/gravity2): same name, same signature, computed on a B200."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

DistMat2Tree = _ref(_d.DistMat2Tree)
