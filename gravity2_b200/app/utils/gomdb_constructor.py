"""Mirror of app/utils/gomdb_constructor.py."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

GOMDB_Constructor = _ref(_d.GOMDB_Constructor)
