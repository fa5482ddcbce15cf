"""Mirror of app/utils/gom_signature_table_constructor.py (4-arg twin)."""
from ...dropin import GOMSignatureTable_Constructor, generate_gom_sigs  # noqa: F401
