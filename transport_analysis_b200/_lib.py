"""ctypes binding of ``libta_b200.so`` (C ABI: ``include/ta_b200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is present, constructing a :class:`Context` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint64, c_void_p

import numpy as np

TA_DTYPE_F32, TA_DTYPE_F64 = 0, 1
TA_PRECISION_FP64, TA_PRECISION_FP32 = 0, 1
TA_LAYOUT_ATOM_MAJOR, TA_LAYOUT_LAG_MAJOR = 0, 1
TA_NCCL_ID_BYTES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libta_b200.so")

# every symbol include/ta_b200.h declares: (restype, argtypes)
_SIGNATURES = {
    "ta_version": (c_int, []),
    "ta_device_count": (c_int, [POINTER(c_int)]),
    "ta_last_error": (c_char_p, [c_void_p]),
    "ta_ctx_create": (c_int, [c_int, POINTER(c_int), POINTER(c_void_p)]),
    "ta_nccl_unique_id": (c_int, [c_void_p]),
    "ta_ctx_create_rank": (c_int, [c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "ta_ctx_destroy": (None, [c_void_p]),
    "ta_host_register": (c_int, [c_void_p, c_uint64]),
    "ta_host_unregister": (c_int, [c_void_p]),
    "ta_stage_begin": (c_int, [c_void_p, c_int64, c_int64, c_int, POINTER(c_int), c_int, c_int,
                               POINTER(c_double), c_int]),
    "ta_stage_slot": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int64)]),
    "ta_stage_commit": (c_int, [c_void_p, c_int64, c_int64]),
    "ta_stage_bulk": (c_int, [c_void_p, POINTER(c_void_p), c_int64, c_int64, c_int64, c_int64, c_int64]),
    "ta_stage_end": (c_int, [c_void_p]),
    "ta_vacf_fft": (c_int, [c_void_p, POINTER(c_double)]),
    "ta_vacf_windowed": (c_int, [c_void_p, POINTER(c_double)]),
    "ta_helfand": (c_int, [c_void_p, POINTER(c_double), c_double, c_double, POINTER(c_double)]),
    "ta_helfand_fft": (c_int, [c_void_p, POINTER(c_double), c_double, c_double, POINTER(c_double)]),
    "ta_fetch_by_particle": (c_int, [c_void_p, c_int64, c_int64, c_int, POINTER(c_double)]),
    "ta_green_kubo": (c_int, [c_void_p, POINTER(c_double), c_int64, c_int64, c_int64, c_double, POINTER(c_double),
                              POINTER(c_double), POINTER(c_double)]),
    "ta_timer_begin": (c_int, [c_void_p]),
    "ta_timer_end": (c_int, [c_void_p, POINTER(c_float)]),
    "ta_last_kernel_ms": (c_int, [c_void_p, POINTER(c_float)]),
    "ta_probe_fp64": (c_int, [c_void_p, POINTER(c_double)]),
    "ta_probe_h2d": (c_int, [c_void_p, c_uint64, POINTER(c_double)]),
    "ta_k1_uses_tmem": (c_int, [c_void_p]),
    "ta_flush_l2": (c_int, [c_void_p]),
    "ta_launch_count": (c_int64, [c_void_p]),
    "ta_helfand_fft_refined": (c_int64, [c_void_p]),
    "ta_fft_plan_info": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                 POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


TA_ERR_UNSUPPORTED = -4   # include/ta_b200.h


class BackendError(RuntimeError):
    """A libta_b200 call failed (the message carries ta_last_error())."""


class UnsupportedError(BackendError):
    """TA_ERR_UNSUPPORTED: the requested route does not serve this problem size / precision."""


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Load libta_b200.so and declare its prototypes.  Raises if it is missing
    (build it with ``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("TA_B200_LIB") or LIB_PATH   # TA_B200_LIB: an experimental build of the same ABI
    if not os.path.exists(path):
        raise BackendError(
            f"{path} not found: the CUDA backend has not been built "
            "(run __graft_entry__.build()); there is no CPU fallback"
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def device_count() -> int:
    n = c_int(0)
    load_library().ta_device_count(ctypes.byref(n))
    return n.value


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = ctypes.create_string_buffer(TA_NCCL_ID_BYTES)
    rc = lib.ta_nccl_unique_id(buf)
    if rc != 0:
        raise BackendError(f"ta_nccl_unique_id failed ({rc}): {lib.ta_last_error(None).decode()}")
    return buf.raw


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_double))


class Context:
    """One analysis context = one ``ta_ctx`` (device memory, streams, NCCL)."""

    def __init__(self, devices=None, rank: int | None = None, nranks: int = 1, nccl_id: bytes | None = None):
        self._lib = load_library()
        self._h = c_void_p()
        if rank is not None:
            dev = 0 if devices is None else int(devices[0] if np.ndim(devices) else devices)
            idbuf = ctypes.create_string_buffer(nccl_id, TA_NCCL_ID_BYTES) if nccl_id else None
            rc = self._lib.ta_ctx_create_rank(dev, int(rank), int(nranks), idbuf, ctypes.byref(self._h))
        else:
            if devices is None:
                devices = [0]
            devs = (c_int * len(devices))(*[int(d) for d in devices])
            rc = self._lib.ta_ctx_create(len(devices), devs, ctypes.byref(self._h))
        if rc != 0:
            msg = self._lib.ta_last_error(None).decode()
            self._h = c_void_p()
            raise BackendError(f"cannot create a B200 context ({rc}): {msg}")
        self.T = self.N = 0
        self._n_fields = 1
        self._np_dtype = np.float32
        self._keepalive = None

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            cls = UnsupportedError if rc == TA_ERR_UNSUPPORTED else BackendError
            raise cls(f"{what} failed ({rc}): {self._lib.ta_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ta_ctx_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- staging ----------------------------------------------------------
    def stage_begin(self, T, N, dims, src_dtype=np.float32, n_fields=1, masses=None, precision="fp64"):
        dims = [int(d) for d in dims]
        cd = (c_int * len(dims))(*dims)
        np_dtype = np.dtype(src_dtype)
        if np_dtype == np.float32:
            code = TA_DTYPE_F32
        elif np_dtype == np.float64:
            code = TA_DTYPE_F64
        else:
            raise ValueError(f"unsupported source dtype {np_dtype}")
        prec = {"fp64": TA_PRECISION_FP64, "fp32": TA_PRECISION_FP32}[precision]
        mptr = None
        if masses is not None:
            self._masses = np.ascontiguousarray(masses, dtype=np.float64)
            mptr = _dptr(self._masses)
        self._check(
            self._lib.ta_stage_begin(self._h, int(T), int(N), len(dims), cd, code, int(n_fields), mptr, prec),
            "ta_stage_begin",
        )
        self.T, self.N, self._n_fields, self._np_dtype = int(T), int(N), int(n_fields), np_dtype

    def stage_slot(self) -> np.ndarray:
        """Pinned slab as an array [capacity, n_fields, N, 3] of the source dtype."""
        ptr, cap = c_void_p(), c_int64()
        self._check(self._lib.ta_stage_slot(self._h, ctypes.byref(ptr), ctypes.byref(cap)), "ta_stage_slot")
        shape = (cap.value, self._n_fields, self.N, 3)
        nbytes = int(np.prod(shape)) * self._np_dtype.itemsize
        buf = (ctypes.c_char * nbytes).from_address(ptr.value)
        return np.frombuffer(buf, dtype=self._np_dtype).reshape(shape)

    def stage_commit(self, frame0: int, nframes: int):
        self._check(self._lib.ta_stage_commit(self._h, int(frame0), int(nframes)), "ta_stage_commit")

    def stage_bulk(self, fields, atom_first=0, frame_first=0, frame_step=1, nframes=None):
        """fields: list of C-contiguous [F, A, 3] arrays of the source dtype."""
        arrs = []
        for a in fields:
            if a.dtype != self._np_dtype or not a.flags.c_contiguous or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("bulk fields must be C-contiguous [frames, atoms, 3] arrays of the staged dtype")
            arrs.append(a)
        src_atoms = arrs[0].shape[1]
        nframes = self.T if nframes is None else nframes
        last = frame_first + (nframes - 1) * frame_step
        if last >= arrs[0].shape[0]:
            raise ValueError("frame window exceeds the source array")
        ptrs = (c_void_p * 2)(*([a.ctypes.data for a in arrs] + [None] * (2 - len(arrs))))
        self._keepalive = arrs
        self._check(
            self._lib.ta_stage_bulk(self._h, ptrs, int(src_atoms), int(atom_first), int(frame_first),
                                    int(frame_step), int(nframes)),
            "ta_stage_bulk",
        )

    def stage_end(self):
        self._check(self._lib.ta_stage_end(self._h), "ta_stage_end")
        self._keepalive = None

    # -- compute ----------------------------------------------------------
    def vacf_fft(self) -> np.ndarray:
        ts = np.empty(self.T, dtype=np.float64)
        self._check(self._lib.ta_vacf_fft(self._h, _dptr(ts)), "ta_vacf_fft")
        self._keepalive = None            # the call has waited for every staging copy
        return ts

    def vacf_windowed(self) -> np.ndarray:
        ts = np.empty(self.T, dtype=np.float64)
        self._check(self._lib.ta_vacf_windowed(self._h, _dptr(ts)), "ta_vacf_windowed")
        self._keepalive = None
        return ts

    def helfand(self, volumes, boltzmann: float, temp_avg: float, fft: bool = False) -> np.ndarray:
        """``fft=False``: the direct windowed MSD (kernel K3).  ``fft=True``: the O(T log T) route S1 - 2 S2 with exact
        re-evaluation of the lags where the difference cancels (kernels K1 + K5 + K6).  Both meet the 1e-10 bar."""
        vol = np.ascontiguousarray(volumes, dtype=np.float64)
        if vol.shape != (self.T,):
            raise ValueError("volumes must have one entry per analysed frame")
        ts = np.empty(self.T, dtype=np.float64)
        fn, name = (self._lib.ta_helfand_fft, "ta_helfand_fft") if fft else (self._lib.ta_helfand, "ta_helfand")
        self._check(fn(self._h, _dptr(vol), float(boltzmann), float(temp_avg), _dptr(ts)), name)
        self._keepalive = None
        return ts

    def fetch_by_particle(self, atom0=0, natoms=None, lag_major_copy=False) -> np.ndarray:
        """Per-particle results as an array of shape (T, natoms), the reference's
        shape (velocityautocorr.py:145-147).  By default it is a transposed view
        of the atom-major device buffer (no extra pass); ``lag_major_copy=True``
        has the device produce the C-contiguous lag-major layout."""
        natoms = self.N - atom0 if natoms is None else natoms
        if lag_major_copy:
            out = np.empty((self.T, natoms), dtype=np.float64)
            self._check(self._lib.ta_fetch_by_particle(self._h, int(atom0), int(natoms), TA_LAYOUT_LAG_MAJOR,
                                                       _dptr(out)), "ta_fetch_by_particle")
            return out
        out = np.empty((natoms, self.T), dtype=np.float64)
        self._check(self._lib.ta_fetch_by_particle(self._h, int(atom0), int(natoms), TA_LAYOUT_ATOM_MAJOR,
                                                   _dptr(out)), "ta_fetch_by_particle")
        return out.T

    def green_kubo(self, times, start=0, stop=None, step=1, initial=0.0, running=False):
        """Trapezoid integral, least-squares slope and (optionally) running integral of the device-resident atom-mean
        timeseries over ``times[start:stop:step]`` (kernel K7).  Returns ``(integral, slope, running | None)``."""
        t = np.ascontiguousarray(times, dtype=np.float64)
        if t.shape != (self.T,):
            raise ValueError("times must have one entry per analysed frame")
        stop = self.T if stop is None else int(stop)
        n = len(range(int(start), stop, int(step)))
        integ, slope = c_double(), c_double()
        run = np.empty(n, dtype=np.float64) if running else None
        self._check(self._lib.ta_green_kubo(self._h, _dptr(t), int(start), stop, int(step), float(initial),
                                            ctypes.byref(integ), _dptr(run) if running else None, ctypes.byref(slope)),
                    "ta_green_kubo")
        return float(integ.value), float(slope.value), run

    # -- timing / introspection -------------------------------------------
    def timer_begin(self):
        self._check(self._lib.ta_timer_begin(self._h), "ta_timer_begin")

    def timer_end(self) -> float:
        ms = c_float()
        self._check(self._lib.ta_timer_end(self._h, ctypes.byref(ms)), "ta_timer_end")
        return float(ms.value)

    def last_kernel_ms(self) -> float:
        ms = c_float()
        self._check(self._lib.ta_last_kernel_ms(self._h, ctypes.byref(ms)), "ta_last_kernel_ms")
        return float(ms.value)

    def probe_fp64(self) -> float:
        """Measured FP64 FMA rate of the first device of this context, TFLOP/s."""
        v = c_double()
        self._check(self._lib.ta_probe_fp64(self._h, ctypes.byref(v)), "ta_probe_fp64")
        return float(v.value)

    def probe_h2d(self, nbytes: int = 1 << 30) -> float:
        """Rate of a contiguous pinned host -> device copy on the first device of this context, GB/s."""
        v = c_double()
        self._check(self._lib.ta_probe_h2d(self._h, c_uint64(int(nbytes)), ctypes.byref(v)), "ta_probe_h2d")
        return float(v.value)

    def flush_l2(self):
        self._check(self._lib.ta_flush_l2(self._h), "ta_flush_l2")

    def launch_count(self) -> int:
        return int(self._lib.ta_launch_count(self._h))

    def helfand_fft_refined(self) -> int:
        """(particle, lag) pairs the last FFT-route Helfand call evaluated exactly; -1 if the direct kernel did it all."""
        return int(self._lib.ta_helfand_fft_refined(self._h))

    def fft_plan_info(self) -> dict:
        H, npz, thr, smem, grid = c_int(), c_int(), c_int(), c_int(), c_int()
        rad = (c_int * 12)()
        self._lib.ta_fft_plan_info(self._h, ctypes.byref(H), ctypes.byref(npz), rad, ctypes.byref(thr),
                                   ctypes.byref(smem), ctypes.byref(grid))
        return {"H": H.value, "radices": list(rad)[: npz.value], "threads": thr.value,
                "smem_bytes": smem.value, "grid": grid.value, "tmem": bool(self._lib.ta_k1_uses_tmem(self._h))}


def host_register(arr: np.ndarray):
    lib = load_library()
    rc = lib.ta_host_register(c_void_p(arr.ctypes.data), c_uint64(arr.nbytes))
    if rc != 0:
        raise BackendError(f"ta_host_register failed ({rc}): {lib.ta_last_error(None).decode()}")


def host_unregister(arr: np.ndarray):
    lib = load_library()
    lib.ta_host_unregister(c_void_p(arr.ctypes.data))


# Page-locked registrations made on behalf of the analysis classes (FrameStager.try_bulk): one per array, keyed by
# address, dropped when the array is garbage-collected.  Arrays the caller registered itself are left alone.
_PINNED: dict = {}


def is_pinned(arr: np.ndarray) -> bool:
    ent = _PINNED.get(arr.ctypes.data)
    return ent is not None and ent >= arr.nbytes


def _unpin_address(addr: int):
    if _PINNED.pop(addr, None) is not None and _lib is not None:
        _lib.ta_host_unregister(c_void_p(addr))


def pin_array(arr: np.ndarray) -> bool:
    """Page-lock ``arr`` (cudaHostRegister) unless that has been done already; returns whether the array is pinned.
    A failure (locked-memory limit, memory the driver cannot pin, a range the caller has already registered) is not
    an error: copies from pageable memory work, through the driver's staging buffers."""
    import weakref

    if is_pinned(arr):
        return True
    lib = load_library()
    addr = arr.ctypes.data
    rc = lib.ta_host_register(c_void_p(addr), c_uint64(arr.nbytes))
    if rc != 0:
        # already registered by the caller (bench.py, a user who pins their own buffers) counts as pinned
        return b"already" in (lib.ta_last_error(None) or b"")
    _PINNED[addr] = arr.nbytes
    owner = arr
    while isinstance(getattr(owner, "base", None), np.ndarray):
        owner = owner.base            # the object whose death frees the memory
    try:
        weakref.finalize(owner, _unpin_address, addr)
    except TypeError:
        pass                          # not weak-referenceable: stays pinned for the life of the process
    return True
