"""CPU oracle for the MVAE training-step hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is on the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
leg may import it, and there only as the checker / the timed CPU baseline.

Parity status: the reference (mhw32/multimodal-vae-public) ships no tests and no
golden vectors, so the oracle is pinned against outputs of the reference itself,
imported unmodified in the build container by ``tests/golden/make_golden.py``
(fixtures committed under ``tests/golden/``).
"""
