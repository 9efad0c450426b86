"""`core` namespace shim: core.loss / core.metric -> B200 path; everything else -> the next `core`
directory on sys.path (the reference checkout)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.append(_root)
__path__ = [_here]
for _p in sys.path:
    _cand = os.path.join(_p or '.', 'core')
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.abspath(_cand) not in __path__:
        __path__.append(os.path.abspath(_cand))
