"""CPU oracle for the Friture spectral hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline -- never on the
product path (``friture_b200`` fails loudly when its CUDA library is missing and has
no CPU fallback).

Parity status: PINNED.  The restatement in :mod:`oracle.friture_oracle` is validated
against the unmodified reference imported from ``/root/reference`` (see
``oracle/ref_import.py`` and ``tests/test_oracle_vs_reference.py``, which run in the
build container) and against golden vectors generated from that reference by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (these travel to the
GPU box, where ``/root/reference`` does not exist).
"""
