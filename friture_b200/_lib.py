"""ctypes binding of ``libfrt_b200.so`` (the C ABI declared in ``include/frt.h``)."""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libfrt_b200.so"

FRT_OK = 0
FRT_EINVAL = -1
FRT_ECUDA = -2
FRT_ENOMEM = -3
FRT_ESTATE = -4

STFT_POWER = 0
STFT_LOGPOWER = 1


class FrtError(RuntimeError):
    """Raised for every non-zero return code of the C ABI (bad argument -> ValueError subclass)."""

    def __init__(self, code, message):
        super().__init__("frt error %d: %s" % (code, message))
        self.code = code


class FrtValueError(FrtError, ValueError):
    pass


def lib_path() -> str:
    # FRT_B200_LIB: tuning knob to A/B another build of the same library
    return os.environ.get("FRT_B200_LIB") or os.path.join(_HERE, _LIB_NAME)


# every exported symbol with (restype, argtypes); tests check that the built library exports
# exactly the functions declared in include/frt.h
_f32p = POINTER(c_float)
SIGNATURES = {
    "frt_version": (c_int, []),
    "frt_create": (c_int, [c_int, POINTER(c_void_p)]),
    "frt_destroy": (c_int, [c_void_p]),
    "frt_last_error": (c_char_p, [c_void_p]),
    "frt_device_sm_count": (c_int, [c_void_p]),
    "frt_launch_count": (c_int64, [c_void_p]),
    "frt_host_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "frt_host_free": (c_int, [c_void_p, c_void_p]),
    "frt_peer_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "frt_peer_free": (c_int, [c_void_p, c_void_p]),
    "frt_peer_export": (c_int, [c_void_p, c_void_p, c_void_p]),
    "frt_peer_import": (c_int, [c_void_p, c_void_p, POINTER(c_void_p)]),
    "frt_peer_close": (c_int, [c_void_p, c_void_p]),
    "frt_peer_copy": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "frt_peer_push": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_size_t, c_int, c_void_p]),
    "frt_stft_plan": (c_int, [c_void_p, c_int]),
    "frt_stft_window": (c_int, [c_void_p, c_void_p]),
    "frt_stft_process": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p,
                                 c_int64, c_int64, c_int, c_void_p]),
    "frt_stft_process_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int,
                                      c_void_p, c_int]),
    "frt_bank_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "frt_bank_reset": (c_int, [c_void_p]),
    "frt_bank_set_weighting": (c_int, [c_void_p, c_void_p]),
    "frt_bank_schedule": (c_int, [c_int, c_int, c_int64, c_void_p, c_void_p]),
    "frt_bank_schedule2": (c_int, [c_int, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "frt_bank_process": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                                 c_int64, c_int, c_void_p]),
    "frt_bank_process_strided": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64,
                                         c_int, c_void_p]),
    "frt_combined_process_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_int, c_void_p,
                                          c_void_p, c_int, c_int]),
    "frt_bank_state_size": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "frt_bank_get_state": (c_int, [c_void_p, c_void_p, c_void_p]),
    "frt_bank_set_state": (c_int, [c_void_p, c_void_p, c_void_p]),
    "frt_firbank_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "frt_firbank_reset": (c_int, [c_void_p]),
    "frt_firbank_process": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "frt_gcc_plan": (c_int, [c_void_p, c_int]),
    "frt_gcc_phat": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int,
                             c_void_p, c_void_p, c_void_p]),
    "frt_spectrum_reduce": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                    c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "frt_decimate_plan": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "frt_decimate_process": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "frt_display_columns": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                    c_void_p, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p,
                                    c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load ``libfrt_b200.so`` (built in-tree by ``__graft_entry__.build()`` /
    ``make -C friture_b200/csrc``).  Raises if it is missing -- there is no fallback path."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.isfile(path):
            raise FrtError(FRT_ESTATE, "%s not built; run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` or `make -C friture_b200/csrc` (no CPU fallback exists)"
                           % path)
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def _ptr(obj):
    """Raw pointer of a torch tensor / numpy array / int / None."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return c_void_p(obj)
    if hasattr(obj, "data_ptr"):
        return c_void_p(obj.data_ptr())
    if hasattr(obj, "ctypes"):
        return c_void_p(obj.ctypes.data)
    raise TypeError("cannot take a pointer of %r" % type(obj))


class Handle:
    """Owns one ``frt_handle`` (one GPU).  Not re-entrant: one handle <-> one stream at a time,
    like the reference's widgets, which call the hot path from a single thread
    (friture/analyzer.py:193-195)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = c_void_p()
        rc = self._lib.frt_create(int(device), ctypes.byref(self._h))
        if rc != FRT_OK:
            msg = self._lib.frt_last_error(None)
            raise FrtError(rc, msg.decode() if msg else "frt_create failed")
        self.device = int(device)

    def check(self, rc):
        if rc == FRT_OK:
            return
        msg = self._lib.frt_last_error(self._h)
        msg = msg.decode() if msg else "?"
        if rc == FRT_EINVAL:
            raise FrtValueError(rc, msg)
        raise FrtError(rc, msg)

    def check_device(self, tensor):
        """A tensor on another GPU than the handle's would surface as an illegal-address error
        deep inside a kernel; say it plainly instead."""
        idx = getattr(getattr(tensor, "device", None), "index", None)
        if idx is not None and idx != self.device:
            raise FrtValueError(FRT_EINVAL, "tensor lives on cuda:%d, this handle is bound to cuda:%d"
                                % (idx, self.device))

    def call(self, name, *args):
        self.check(getattr(self._lib, name)(self._h, *args))

    @property
    def launch_count(self) -> int:
        return int(self._lib.frt_launch_count(self._h))

    @property
    def sm_count(self) -> int:
        return int(self._lib.frt_device_sm_count(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.frt_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_handles = {}


def default_handle(device: int | None = None) -> Handle:
    """Process-wide handle per device (the drop-in classes share it, as the reference's widgets
    share one process)."""
    if device is None:
        import torch
        if not torch.cuda.is_available():
            raise FrtError(FRT_ECUDA, "no CUDA device available; friture_b200 has no CPU fallback")
        device = torch.cuda.current_device()
    h = _default_handles.get(device)
    if h is None:
        h = Handle(device)
        _default_handles[device] = h
    return h


def current_stream_ptr(device=None):
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
