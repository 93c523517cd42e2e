"""ctypes binding of the C ABI in ``include/dpdfnet_b200.h``.

The Python layer is plumbing only: it packs a checkpoint, hands raw device (or host) pointers to
the shared library and raises the reference's exception types on failure.  There is no CPU or
PyTorch fallback - if the library is missing or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Dict, Mapping, Optional, Sequence, Union

import numpy as np

from .spec import ModelSpec, get_spec
from . import weights as _weights

FLAG_WARMUP = 1
FLAG_ZERO_FEAT = 2
FLAG_ZERO_SPEC = 8

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libdpdfnet_b200.so"
_lib = None


class _Spec(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("sample_rate", ctypes.c_int32), ("win", ctypes.c_int32),
                ("hop", ctypes.c_int32), ("freq_bins", ctypes.c_int32), ("n_blocks", ctypes.c_int32),
                ("hr48", ctypes.c_int32), ("fe_feat", ctypes.c_int32), ("fe", ctypes.c_int32 * 4),
                ("erb_strides", ctypes.c_int32 * 3), ("dec_up", ctypes.c_int32 * 3),
                ("erb_widths", ctypes.c_int32 * 32), ("state_size", ctypes.c_int32)]


C_API = {
    # name: (restype, argtypes)
    "dpdf_create": (ctypes.c_int, [ctypes.POINTER(_Spec), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32,
                                   ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    "dpdf_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dpdf_reset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_step_spec": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_step_pcm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_run_pcm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_prime_pcm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                      ctypes.c_void_p]),
    "dpdf_step_spec_host": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32]),
    "dpdf_step_pcm_host": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32]),
    "dpdf_run_pcm_host": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int32, ctypes.c_int32]),
    "dpdf_submit_pcm_host": (ctypes.c_int, [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]),
    "dpdf_wait": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "dpdf_state_size": (ctypes.c_int, [ctypes.c_void_p]),
    "dpdf_state_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_state_import": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "dpdf_debug_tensor": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_size_t)]),
    "dpdf_kernel_launches": (ctypes.c_int, [ctypes.c_void_p]),
    "dpdf_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32]),
    "dpdf_time_kernels": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]),
    "dpdf_resampler_create": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p)]),
    "dpdf_resampler_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "dpdf_resampler_reset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dpdf_resampler_pending": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]),
    "dpdf_resampler_process": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p,
                                              ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64),
                                              ctypes.c_void_p]),
    "dpdf_poll_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "dpdf_last_error": (ctypes.c_char_p, []),
    "dpdf_version": (ctypes.c_char_p, []),
}


def load_library(path: Optional[Union[str, Path]] = None):
    """Load ``libdpdfnet_b200.so`` and declare every entry point of the header."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else Path(os.environ.get("DPDFNET_B200_LIB", _LIB_PATH))
    if not p.is_file():
        raise FileNotFoundError(
            f"CUDA engine library not found: {p}. Build it with `python -m dpdfnet_b200.build` "
            "(there is no CPU fallback).")
    lib = ctypes.CDLL(str(p))
    for name, (res, args) in C_API.items():
        fn = getattr(lib, name)          # AttributeError if the ABI is incomplete
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def _raise(lib, rc: int):
    msg = lib.dpdf_last_error().decode(errors="replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -2:
        raise ValueError(f"weights: {msg}")
    if rc == -4:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def c_spec(spec: ModelSpec) -> _Spec:
    s = _Spec()
    s.abi_version = 1
    s.sample_rate, s.win, s.hop, s.freq_bins = spec.sample_rate, spec.win, spec.hop, spec.freq_bins
    s.n_blocks, s.hr48, s.fe_feat = spec.n_blocks, int(spec.hr48), spec.fe_feat
    s.fe = (ctypes.c_int32 * 4)(*spec.fe)
    s.erb_strides = (ctypes.c_int32 * 3)(*spec.erb_strides)
    s.dec_up = (ctypes.c_int32 * 3)(*spec.dec_up)
    s.erb_widths = (ctypes.c_int32 * 32)(*spec.erb_widths)
    s.state_size = spec.state_size
    return s


def _ptr(t) -> int:
    """Raw pointer of a torch tensor / numpy array / None."""
    if t is None:
        return 0
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class Engine:
    """Batched streaming DPDFNet engine on one B200.

    ``checkpoint`` is a reference ``state_dict`` (offline naming, numpy or torch tensors), a packed
    blob (``bytes``) or ``None`` for the seeded random stand-in (``seed``).
    """

    def __init__(self, model: Union[str, ModelSpec], checkpoint=None, *, max_streams: int = 1, device: int = 0,
                 seed: int = 0):
        self.spec = get_spec(model) if isinstance(model, str) else model
        self._lib = load_library()
        if isinstance(checkpoint, (bytes, bytearray, memoryview)):
            blob = bytes(checkpoint)
        else:
            if checkpoint is None:
                checkpoint = _weights.random_checkpoint(self.spec, seed)
            sd = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in checkpoint.items()}
            blob = _weights.pack_checkpoint(self.spec, sd)
        self._handle = ctypes.c_void_p()
        cs = c_spec(self.spec)
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self._lib.dpdf_create(ctypes.byref(cs), buf, len(blob), int(max_streams), int(device), ctypes.byref(self._handle))
        if rc != 0:
            self._handle = None
            _raise(self._lib, rc)
        self.max_streams = int(max_streams)
        self.device = int(device)

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None):
            self._lib.dpdf_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            _raise(self._lib, rc)

    def _stream(self) -> int:
        """torch's current stream ON THE ENGINE'S DEVICE (not on whatever device is current)."""
        import torch
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_int(self, a, B):
        """slot ids / flags given as None, torch cuda int32 tensor, or host sequence."""
        if a is None:
            return None
        import torch
        if isinstance(a, torch.Tensor):
            t = a.to(device=f"cuda:{self.device}", dtype=torch.int32).contiguous()
        else:
            t = torch.as_tensor(np.asarray(a, dtype=np.int32), device=f"cuda:{self.device}")
        if t.numel() != B:
            raise ValueError(f"expected {B} entries, got {t.numel()}")
        return t

    # ----- device-tensor entry points (torch tensors are containers only) --
    def reset(self, slots: Optional[Sequence[int]] = None):
        """``None`` resets every slot; a (possibly empty) list resets exactly those slots."""
        if slots is None:
            self._check(self._lib.dpdf_reset(self._handle, None, 0, self._stream()))
            return
        arr = np.ascontiguousarray(np.asarray(slots, dtype=np.int32).reshape(-1))
        if arr.size == 0:                # the C ABI reads n <= 0 as "all slots": an empty list must reset nothing
            return
        self._check(self._lib.dpdf_reset(self._handle, arr.ctypes.data, arr.size, self._stream()))
        import torch
        torch.cuda.current_stream(self.device).synchronize()       # host slot list must outlive the async copy

    def step_spec(self, spec_in, slot_ids=None, flags=None, out=None):
        import torch
        B = spec_in.shape[0]
        assert spec_in.is_cuda and spec_in.dtype == torch.float32 and spec_in.is_contiguous()
        assert tuple(spec_in.shape[1:]) == (self.spec.freq_bins, 2)
        out = torch.empty_like(spec_in) if out is None else out
        s, f = self._dev_int(slot_ids, B), self._dev_int(flags, B)
        self._check(self._lib.dpdf_step_spec(self._handle, _ptr(spec_in), _ptr(out), _ptr(s), _ptr(f), B, self._stream()))
        return out

    def step_pcm(self, pcm, slot_ids=None, flags=None, out=None):
        import torch
        B = pcm.shape[0]
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.shape[1] == self.spec.hop and pcm.stride(1) == 1
        out = torch.empty((B, self.spec.hop), device=pcm.device, dtype=torch.float32) if out is None else out
        s, f = self._dev_int(slot_ids, B), self._dev_int(flags, B)
        self._check(self._lib.dpdf_step_pcm(self._handle, _ptr(pcm), pcm.stride(0), _ptr(out), out.stride(0),
                                            _ptr(s), _ptr(f), B, self._stream()))
        return out

    def run_pcm(self, pcm, slot_ids=None, flags=None, out=None):
        """T = pcm.shape[1] // hop consecutive hops, one CUDA-graph replay each."""
        import torch
        B, n = pcm.shape
        T = n // self.spec.hop
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.stride(1) == 1 and T > 0
        out = torch.zeros_like(pcm) if out is None else out
        s, f = self._dev_int(slot_ids, B), self._dev_int(flags, B)
        self._check(self._lib.dpdf_run_pcm(self._handle, _ptr(pcm), pcm.stride(0), _ptr(out), out.stride(0),
                                           _ptr(s), _ptr(f), B, T, self._stream()))
        return out

    def prime_pcm(self, pcm, slot_ids=None):
        B = pcm.shape[0]
        assert pcm.is_cuda and pcm.shape[1] >= self.spec.hop and pcm.stride(1) == 1
        s = self._dev_int(slot_ids, B)
        self._check(self._lib.dpdf_prime_pcm(self._handle, _ptr(pcm), pcm.stride(0), _ptr(s), B, self._stream()))

    # ----- host-buffer entry points (what a reference-side binding calls) --
    @staticmethod
    def _host_int(a, B):
        if a is None:
            return None
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))
        if arr.size != B:
            raise ValueError(f"expected {B} entries, got {arr.size}")
        return arr

    def step_spec_host(self, spec_in: np.ndarray, slot_ids=None, flags=None) -> np.ndarray:
        x = np.ascontiguousarray(spec_in, dtype=np.float32)
        if x.ndim != 3 or x.shape[1:] != (self.spec.freq_bins, 2):
            raise ValueError(f"spec must be [B, {self.spec.freq_bins}, 2], got {x.shape}")
        B = x.shape[0]
        out = np.empty_like(x)
        s, f = self._host_int(slot_ids, B), self._host_int(flags, B)
        self._check(self._lib.dpdf_step_spec_host(self._handle, _ptr(x), _ptr(out), _ptr(s), _ptr(f), B))
        return out

    def step_pcm_host(self, pcm: np.ndarray, slot_ids=None, flags=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.spec.hop:
            raise ValueError(f"pcm must be [B, {self.spec.hop}], got {x.shape}")
        B = x.shape[0]
        if out is None:
            out = np.empty_like(x)
        elif out.shape != x.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array shaped like pcm")
        s, f = self._host_int(slot_ids, B), self._host_int(flags, B)
        self._check(self._lib.dpdf_step_pcm_host(self._handle, _ptr(x), _ptr(out), _ptr(s), _ptr(f), B))
        return out

    def submit_pcm_host(self, pcm: np.ndarray, out: np.ndarray, slot_ids=None, flags=None) -> int:
        """Pipelined form of ``step_pcm_host``: returns a ticket at once; ``wait(ticket)`` delivers ``out``.  ``pcm`` and
        ``out`` must be C-contiguous float32 [B, hop] arrays that stay alive (ideally pinned) until the wait."""
        if pcm.dtype != np.float32 or not pcm.flags.c_contiguous or pcm.ndim != 2 or pcm.shape[1] != self.spec.hop:
            raise ValueError(f"pcm must be a C-contiguous float32 [B, {self.spec.hop}] array")
        if out.shape != pcm.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array shaped like pcm")
        B = pcm.shape[0]
        s, f = self._host_int(slot_ids, B), self._host_int(flags, B)
        t = ctypes.c_int64()
        self._check(self._lib.dpdf_submit_pcm_host(self._handle, _ptr(pcm), _ptr(out), _ptr(s), _ptr(f), B, ctypes.byref(t)))
        return int(t.value)

    def wait(self, ticket: int) -> None:
        self._check(self._lib.dpdf_wait(self._handle, int(ticket)))

    def run_pcm_host(self, pcm: np.ndarray, slot_ids=None) -> np.ndarray:
        x = np.ascontiguousarray(pcm, dtype=np.float32)
        B, n = x.shape
        T = n // self.spec.hop
        if T <= 0 or n != T * self.spec.hop:
            raise ValueError("pcm length must be a positive multiple of the hop size")
        out = np.empty_like(x)
        s = self._host_int(slot_ids, B)
        self._check(self._lib.dpdf_run_pcm_host(self._handle, _ptr(x), _ptr(out), _ptr(s), B, T))
        return out

    def prime_pcm_host(self, pcm: np.ndarray, slot_ids=None):
        import torch
        x = torch.as_tensor(np.ascontiguousarray(pcm, dtype=np.float32), device=f"cuda:{self.device}")
        self.prime_pcm(x, slot_ids)
        torch.cuda.current_stream(self.device).synchronize()

    # ----- state ----------------------------------------------------------
    @property
    def state_size(self) -> int:
        return self._lib.dpdf_state_size(self._handle)

    def state_export(self, slot: int) -> np.ndarray:
        out = np.empty(self.state_size, np.float32)
        self._check(self._lib.dpdf_state_export(self._handle, int(slot), out.ctypes.data))
        return out

    def state_import(self, slot: int, flat: np.ndarray):
        a = np.ascontiguousarray(flat, dtype=np.float32).reshape(-1)
        if a.size != self.state_size:
            raise ValueError(f"state size mismatch: expected {self.state_size}, got {a.size}")
        self._check(self._lib.dpdf_state_import(self._handle, int(slot), a.ctypes.data))

    # ----- introspection ----------------------------------------------------
    def debug_tensor(self, name: str, B: int) -> np.ndarray:
        n = ctypes.c_size_t()
        self._check(self._lib.dpdf_debug_tensor(self._handle, name.encode(), None, 0, ctypes.byref(n)))
        out = np.empty((B, n.value), np.float32)
        self._check(self._lib.dpdf_debug_tensor(self._handle, name.encode(), out.ctypes.data, out.size, ctypes.byref(n)))
        return out

    def poll_error(self, synchronize: bool = True):
        """Raise ``RuntimeError`` if a kernel of an earlier hop flagged an invalid hop (see dpdf_poll_error)."""
        self._check(self._lib.dpdf_poll_error(self._handle, int(bool(synchronize))))

    @property
    def kernel_launches(self) -> int:
        return self._lib.dpdf_kernel_launches(self._handle)

    def set_option(self, key: str, value: int):
        self._check(self._lib.dpdf_set_option(self._handle, key.encode(), int(value)))

    def time_kernels(self, B: int, iters: int = 5) -> Dict[str, float]:
        """Average device milliseconds per kernel of one hop (un-graphed, CUDA events)."""
        ms = (ctypes.c_float * 256)()
        names = (ctypes.c_char_p * 256)()
        n = ctypes.c_int32()
        self._check(self._lib.dpdf_time_kernels(self._handle, B, iters, ms, names, 256, ctypes.byref(n)))
        out: Dict[str, float] = {}
        for i in range(n.value):
            k = names[i].decode()
            out[k] = out.get(k, 0.0) + float(ms[i])
        return out
