"""CPU oracle for the fusion-loss / fusion-metric hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never as a compute path behind the
drop-in API.  The product path (``multi-modal-image-fusion_b200``) raises if its
CUDA library is missing; it never routes through this package.

The oracle restates, op for op, the algorithm of the reference repository's
``core/loss.py`` and ``core/metric.py`` (plus the ``eval_metrics`` composition
of ``eval.py:29-75``) with plain torch CPU ops, dtype-generic so it can be run
in float32 (the reference's arithmetic) and in float64 (the "truth" the
dual parity gate of SURVEY.md §8(c) needs).

Parity pin: ``tests/golden/make_golden.py`` imports the *real* reference from
``/root/reference`` in the build container and stores its outputs on seeded
inputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this oracle against those vectors (bit-exact for histogram counts, <=1e-6
relative for floating point, 0 for most).  The reference itself ships no tests
or golden vectors for this path, so these generated fixtures
are the pin.
"""
from . import fusion_loss, fusion_metric  # noqa: F401
