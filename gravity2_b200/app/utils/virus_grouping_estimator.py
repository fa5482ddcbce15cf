"""Mirror of app/utils/virus_grouping_estimator.py:74-89 (linkage on the device)."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

VirusGrouping_Estimator = _ref(_d.VirusGrouping_Estimator)
