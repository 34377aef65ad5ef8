"""``from networks.Transception import Transception`` resolves here exactly as in the reference tree; the implementation
lives in ``transception_b200.transception``."""
from transception_b200.transception import *  # noqa: F401,F403
from transception_b200.transception import Transception  # noqa: F401
