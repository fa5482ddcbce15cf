"""Mirror of the numeric half of app/utils/shared_pphmm_graphs.py:59-97."""
from ...dropin import shared_pphmm_ratio, shared_norm_pphmm_ratio  # noqa: F401
