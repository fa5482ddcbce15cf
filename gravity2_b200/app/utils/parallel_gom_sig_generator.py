"""Mirror of app/utils/parallel_gom_sig_generator.py (5-arg GOMSignatureTable_Constructor)."""
from ...dropin import GOMSignatureTable_Constructor, generate_gom_sigs  # noqa: F401
