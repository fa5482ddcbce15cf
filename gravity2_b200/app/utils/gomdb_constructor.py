"""Mirror of app/utils/gomdb_constructor.py."""
from ...dropin import GOMDB_Constructor  # noqa: F401
