"""Import the UNMODIFIED reference (tlecomte/friture) in place from /root/reference.

TEST INFRASTRUCTURE ONLY.  Works only where ``/root/reference`` exists (the build
container); the GPU box has no reference, so nothing that runs there may call this.

``friture.audioproc`` and ``friture.ringbuffer`` import two constants from
``friture.audiobackend`` (audioproc.py:24, ringbuffer.py:25), which drags in
PyQt6/sounddevice/rtmixer (audiobackend.py:24-26; none installed).  A stub module with
just ``SAMPLING_RATE`` / ``FRAMES_PER_BUFFER`` (audiobackend.py:31-32) is pre-seeded in
``sys.modules``; everything else is the reference's own code.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FRITURE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "friture", "audioproc.py"))


def load():
    """Return a namespace with the reference modules of the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True          # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "friture.audiobackend" not in sys.modules:
        stub = types.ModuleType("friture.audiobackend")
        stub.SAMPLING_RATE = 48000
        stub.FRAMES_PER_BUFFER = 512
        sys.modules["friture.audiobackend"] = stub
    import friture.audioproc as audioproc
    import friture.filter as filter_
    import friture.octavefilters as octavefilters
    import friture.generated_filters as generated_filters
    import friture.ringbuffer as ringbuffer
    import friture.signal.lfilter as lfilter
    import friture.signal.decimate as decimate
    import friture.signal.exp_smoothing as exp_smoothing
    import friture.signal.correlation as correlation
    return types.SimpleNamespace(
        audioproc=audioproc, filter=filter_, octavefilters=octavefilters,
        generated_filters=generated_filters, ringbuffer=ringbuffer, lfilter=lfilter,
        decimate=decimate, exp_smoothing=exp_smoothing, correlation=correlation)
