"""``from networks.MSTr import MSTransception`` resolves here exactly as in the reference
(reference ``train_MSTransception.py:12``, ``test.py:14``); the implementation lives in
``transception_b200.mstr``."""
from transception_b200.mstr import *  # noqa: F401,F403
from transception_b200.mstr import MSTransception  # noqa: F401
