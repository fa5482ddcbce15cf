"""Mirror of app/utils/gom_signature_table_constructor.py (4-arg twin)."""
# native failures surface as the reference's SystemExit("FATAL ERROR: ...") here (gravity2_b200/errors.py)
from ... import dropin as _d
from ...errors import reference_errors as _ref

GOMSignatureTable_Constructor = _ref(_d.GOMSignatureTable_Constructor)
generate_gom_sigs = _ref(_d.generate_gom_sigs)
