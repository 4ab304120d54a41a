"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the reference's EEG-CLIP hot path.  Nothing under
``oracle/`` is product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker or the CPU baseline, never as the thing shipped.

Parity status: the reference ships no golden vectors, so the
oracle is pinned against outputs of the *unmodified reference modules* imported
from /root/reference in the build container; the generating script is
``oracle/make_golden.py`` and the committed outputs live in ``tests/golden/``.
"""
