"""transception_b200 — Blackwell-native hot path of TransCeption."""
from .mstr import MSTransception  # noqa: F401
from .transception import Transception  # noqa: F401

__version__ = "0.1.0"
