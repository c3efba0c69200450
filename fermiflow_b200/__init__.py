"""fermiflow_b200 -- B200-native implementation of FermiFlow's per-walker VMC hot path."""
from . import _lib  # noqa: F401
