"""Reference module path ``lib.utils.threading`` (/root/reference/lib/utils/threading.py): the helpers live in ``threads.py``
(named so that nothing in this package can shadow the standard library's ``threading``); this module re-exports them."""
from .threads import *  # noqa: F401,F403
