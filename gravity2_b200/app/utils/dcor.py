"""Mirror of app/utils/dcor.py (dcor only: cent_dist/dcov/dvar are internal to the device kernel)."""
from ...dropin import dcor  # noqa: F401
