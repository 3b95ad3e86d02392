"""ctypes binding of `libb2r.so` (C ABI: include/b2r.h) -- the only road from the Python API to the GPU.

There is NO fallback path: if the shared library is missing, cannot be loaded, or no CUDA device is present,
importing / using this module raises.  (The CPU oracle under `oracle/` is test infrastructure and is never
imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
from enum import Flag, auto

import numpy as np

from . import _abi
from ._abi import (B2R_BG_COLOR, B2R_BG_CUBEMAP, DebugOut, FrameParams, PackedScene, View,  # noqa: F401
                   pack_frame_params, pack_view)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2R_LIB") or os.path.join(HERE, "libb2r.so")  # B2R_LIB: load a tuning variant


class Errors(Flag):
    """Per-face discard reasons printed by the reference's render loop (triangular.py:15-20)."""
    BACK_FACE_CULLING = auto()
    WRONG_MIN_MAX = auto()
    EMPTY_B = auto()
    EMPTY_Z = auto()
    CLIPPED = auto()


_lib = None
_inited_device = None

_SYMBOLS = {
    "b2r_abi_version": (C.c_int, []),
    "b2r_last_error": (C.c_char_p, []),
    "b2r_init": (C.c_int, [C.c_int]),
    "b2r_shutdown": (C.c_int, []),
    "b2r_current_device": (C.c_int, []),
    "b2r_scene_set_silhouette": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "b2r_scene_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2r_scene_destroy": (C.c_int, [C.c_void_p]),
    "b2r_scene_reset_silhouette": (C.c_int, [C.c_void_p]),
    "b2r_scene_device_bytes": (C.c_int64, [C.c_void_p]),
    "b2r_scene_get_silhouette": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "b2r_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "b2r_sync": (C.c_int, []),
    "b2r_last_ticket": (C.c_int64, []),
    "b2r_wait": (C.c_int, [C.c_int64]),
    "b2r_stream": (C.c_void_p, []),
    "b2r_launch_count": (C.c_int64, []),
    "b2r_last_stage_ms": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b2r_set_stage_timing": (C.c_int, [C.c_int]),
    "b2r_window_create": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "b2r_window_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2r_window_close": (C.c_int, [C.c_void_p]),
    "b2r_window_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "b2r_window_destroy": (C.c_int, [C.c_void_p]),
    "b2r_host_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "b2r_host_free": (C.c_int, [C.c_void_p]),
    "b2r_obj_load": (C.c_int, [C.c_char_p, C.c_void_p]),
    "b2r_obj_free": (None, [C.c_void_p]),
}


class ObjArrays(C.Structure):
    """`b2r_obj` of include/b2r.h: what the native OBJ tokenizer hands back."""
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("faces", C.POINTER(C.c_int32)), ("slot_names", C.c_char_p), ("mtllibs", C.c_char_p),
                ("n_vertices", C.c_int32), ("n_uv", C.c_int32), ("n_normals", C.c_int32), ("n_faces", C.c_int32)]


def load_obj(path):
    """Native OBJ tokenizer (host code in libb2r.so, no GPU involved) -> (vertices, uv, normals, faces, slot names,
    mtllib names) with the dtypes / shapes of the reference's loader (core.py:311-315)."""
    lib = load_library()
    o = ObjArrays()
    rc = lib.b2r_obj_load(os.fsencode(path), C.byref(o))
    if rc == 2:
        if os.path.isdir(path):
            raise IsADirectoryError(path)
        raise FileNotFoundError(path)
    if rc == 3:  # the reference's np.array(tokens, dtype=np.int32) raises ValueError on a non-numeric index (core.py:72-74)
        raise ValueError(f"malformed face statement in {path!r}")
    if rc == 4:
        raise MemoryError(path)
    if rc != 0:
        raise RuntimeError(f"b2r_obj_load({path!r}) failed ({rc})")
    try:
        def take(ptr, n, width, dtype):
            return np.ctypeslib.as_array(ptr, shape=(n, width)).astype(dtype, copy=True) if n else None
        vertices = take(o.vertices, o.n_vertices, 4, np.float32)
        uv = take(o.uv, o.n_uv, 3, np.float32)
        normals = take(o.normals, o.n_normals, 3, np.float32)
        faces = (np.ctypeslib.as_array(o.faces, shape=(o.n_faces, 3, 4)).astype(np.int32, copy=True)
                 if o.n_faces else np.zeros((0, 3, 4), np.int32))
        slots = o.slot_names.decode().split("\n")[:-1]
        libs = o.mtllibs.decode().split("\n")[:-1]
    finally:
        lib.b2r_obj_free(C.byref(o))
    if vertices is None:
        vertices = np.zeros((0, 4), np.float32)
    return vertices, uv, normals, faces, slots, libs


def load_library():
    """dlopen libb2r.so and bind every symbol include/b2r.h declares (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m py_numpy_renderer_b200.build` "
                              f"(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        if lib.b2r_abi_version() != _abi.B2R_ABI_VERSION:
            raise ImportError("libb2r.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def _check(rc):
    if rc == _abi.B2R_ERR_INDEX:  # a texture lookup fell outside its map: the reference's fancy indexing raises here
        raise IndexError("b2r: " + (load_library().b2r_last_error() or b"").decode())
    if rc != 0:
        raise RuntimeError("b2r: " + (load_library().b2r_last_error() or b"").decode())


def init(device=None):
    """Bind the library to a CUDA device (default: LOCAL_RANK or 0)."""
    global _inited_device
    lib = load_library()
    explicit = device is not None
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _inited_device is None else _inited_device
    if explicit or _inited_device != device:
        # creates the device's context on first use, afterwards only makes it the calling thread's current one
        _check(lib.b2r_init(int(device)))
        _inited_device = device
    return lib


def sync():
    _check(init().b2r_sync())


_worker = None


def submit(fn, *args, **kw):
    """Run `fn` on the library's single worker thread -> concurrent.futures.Future.  ctypes drops the GIL inside
    libb2r, so the caller can evaluate the cameras of the next batch while this one renders and copies."""
    global _worker
    if _worker is None:
        import sys
        from concurrent.futures import ThreadPoolExecutor
        _worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="b2r")
        # the worker re-acquires the GIL after every library call; with the default 5 ms switch interval it would
        # wait behind the caller's camera maths for longer than a whole batch takes to render
        sys.setswitchinterval(min(sys.getswitchinterval(), 2e-4))
    return _worker.submit(fn, *args, **kw)


def launch_count() -> int:
    return int(init().b2r_launch_count())


def set_stage_timing(flag: bool):
    init().b2r_set_stage_timing(int(flag))


def last_stage_ms() -> dict:
    lib = init()
    names = (C.c_char_p * _abi.B2R_MAX_STAGES)()
    ms = (C.c_float * _abi.B2R_MAX_STAGES)()
    n = lib.b2r_last_stage_ms(names, ms)
    out = {}
    for i in range(n):
        key = names[i].decode()
        out[key] = out.get(key, 0.0) + float(ms[i])
    return out


def stream_ptr() -> int:
    return int(init().b2r_stream() or 0)


def _dev_ptr(t):
    """Address of a device buffer: torch tensor (data_ptr) or a raw int."""
    if t is None:
        return None
    return C.c_void_p(int(t.data_ptr()) if hasattr(t, "data_ptr") else int(t))


class _PinnedPool:
    """Page-locked buffers behind the arrays `Scene.render()` / `render_batch()` hand out when the caller passes no
    `out`: the device-to-host copy of a frame runs at the PCIe rate instead of being staged through the driver's bounce
    buffer, and a buffer returns to the pool when the last view of its array is garbage collected.  At most `limit`
    buffers are outstanding or cached; beyond that ordinary (pageable) arrays are returned."""

    def __init__(self, limit=12, max_bytes=1 << 30):
        self.limit, self.max_bytes = limit, max_bytes
        self.free = {}          # nbytes -> [ptr]
        self.count = 0

    def array(self, shape):
        import weakref
        n = int(np.prod(shape))
        if n <= 0 or n > self.max_bytes:
            return None
        lib = init()
        bucket = self.free.get(n)
        if bucket:
            ptr = bucket.pop()
        else:
            if self.count >= self.limit:
                self._trim()
                if self.count >= self.limit:
                    return None
            p = C.c_void_p()
            if lib.b2r_host_alloc(n, C.byref(p)) != 0:
                return None
            ptr = p.value
            self.count += 1
        raw = (C.c_uint8 * n).from_address(ptr)
        weakref.finalize(raw, self._release, n, ptr)     # runs when the last array viewing `raw` is gone
        return np.ctypeslib.as_array(raw).reshape(shape)

    def _release(self, n, ptr):
        self.free.setdefault(n, []).append(ptr)

    def _trim(self):
        """Drop cached buffers (of sizes nobody asked for lately) to make room."""
        lib = load_library()
        for n in list(self.free):
            while self.free[n]:
                lib.b2r_host_free(C.c_void_p(self.free[n].pop()))
                self.count -= 1
            del self.free[n]


_pinned_pool = _PinnedPool()


class DeviceScene:
    """Device mirror of a list of host `Model`s (+ optional CubeMap)."""

    def __init__(self, models, skybox=None, device=None):
        self.lib = init(device)
        self.packed = PackedScene(models, skybox)
        self.handle = C.c_void_p()
        p = self.packed
        _check(self.lib.b2r_scene_create(C.cast(p.models, C.c_void_p), p.n_models, C.cast(p.textures, C.c_void_p),
                                         p.n_textures, C.cast(p.sky_ptr, C.c_void_p) if p.sky is not None else None,
                                         C.byref(self.handle)))
        self.has_sky = skybox is not None
        self.models_py = list(models)
        self.vertex_base = np.cumsum([0] + [int(m.n_vertices) for m in p.models[:p.n_models]])

    def close(self):
        if self.handle:
            self.lib.b2r_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self):
        return int(self.lib.b2r_scene_device_bytes(self.handle))

    def reset_silhouette(self):
        _check(self.lib.b2r_scene_reset_silhouette(self.handle))

    def silhouette_of(self, model_index):
        cap = 1 << 16
        while True:
            pairs = np.empty((cap, 2), np.int32)
            owner = np.empty(cap, np.int32)
            n = self.lib.b2r_scene_get_silhouette(self.handle, pairs.ctypes.data, owner.ctypes.data, cap)
            if n < 0:
                raise RuntimeError("b2r_scene_get_silhouette failed")
            if n <= cap:
                break
            cap = n
        sel = owner[:n] == model_index
        base = int(self.vertex_base[model_index])
        return {(int(a) - base, int(b) - base) for a, b in pairs[:n][sel]}

    def restore_silhouette(self, per_model_sets):
        """Upload persistent silhouette sets (one set of local (a, b) pairs per model, None / empty = nothing)."""
        pairs = [(int(a) + int(self.vertex_base[i]), int(b) + int(self.vertex_base[i]))
                 for i, sset in enumerate(per_model_sets) if sset for a, b in sset]
        arr = np.ascontiguousarray(np.array(pairs, np.int32).reshape(-1, 2))
        _check(self.lib.b2r_scene_set_silhouette(self.handle, arr.ctypes.data, len(arr)))

    def pack(self, cameras, debug_cameras, light, resolution, system, background, persist_silhouette=False, band=None,
             shading='general'):
        """Host-side evaluation of the per-view constants (the reference's NumPy camera maths) -> (fp, views)."""
        fp = pack_frame_params(light, (int(resolution[0]), int(resolution[1])), background, persist_silhouette, band,
                               shading)
        views = _abi.pack_views(cameras, debug_cameras, system, self.has_sky)
        return fp, views

    def render(self, cameras, debug_cameras, light, resolution, system, background, persist_silhouette=False,
               want_debug=False, out=None, band=None, shading='general'):
        fp, views = self.pack(cameras, debug_cameras, light, resolution, system, background, persist_silhouette, band,
                              shading)
        return self.render_packed(fp, views, want_debug=want_debug, out=out)

    def render_packed(self, fp, views, want_debug=False, out=None, wait=True):
        """-> (frames, info).  frames: uint8 (n, H, W, 3): a new NumPy array, `out` if it is a NumPy array (filled
        synchronously; pinned memory makes the copy fast), or `out` if it is a CUDA tensor / device pointer
        (written asynchronously on the library stream)."""
        n = len(views)
        H, W = fp.height, fp.width
        band = None if (fp.row_begin, fp.row_end) == (0, H) else (fp.row_begin, fp.row_end)
        info = {}
        dbg = None
        on_device = out is not None and not isinstance(out, np.ndarray)
        if want_debug:
            if on_device:
                raise ValueError("debug planes are only returned together with host frames")
            F, M = self.packed.total_faces, self.packed.n_models
            info = dict(face_status=np.zeros((n, max(F, 1)), np.uint8), n_silhouette=np.zeros((n, max(M, 1)), np.int32))
            if want_debug != 'status':  # full planes
                info.update(z=np.empty((n, H, W), np.float64), stencil=np.empty((n, H, W), np.int16),
                            winner=np.empty((n, H, W), np.int32))
            if want_debug == 'overlay':
                info.update(frame_f32=np.empty((n, H, W, 3), np.float32))
            dbg = DebugOut(*[info[k].ctypes.data if k in info else None
                             for k in ('z', 'stencil', 'winner', 'face_status', 'n_silhouette', 'frame_f32')])
        if on_device:
            target, frames = _dev_ptr(out), out
        else:
            if out is not None:
                if out.dtype != np.uint8 or out.shape != (n, H, W, 3) or not out.flags.c_contiguous:
                    raise ValueError("out must be a C-contiguous uint8 array of shape (n, H, W, 3)")
                frames = out
            else:
                frames = _pinned_pool.array((n, H, W, 3))       # page-locked when the pool has room
                if frames is None:
                    frames = np.empty((n, H, W, 3), np.uint8)
                if band is not None:
                    frames[...] = 0
            target = C.c_void_p(frames.ctypes.data)
        mode = 1 if on_device else (0 if (wait or want_debug) else 2)
        _check(self.lib.b2r_render(self.handle, C.byref(fp), C.cast(views, C.c_void_p), n, target,
                                   C.byref(dbg) if dbg is not None else None, mode))
        if mode == 2:
            info['ticket'] = int(self.lib.b2r_last_ticket())
            info['keep'] = (fp, views)
        if want_debug:
            info['face_status'] = info['face_status'][:, :self.packed.total_faces]
            info['n_silhouette'] = info['n_silhouette'][:, :self.packed.n_models]
        return frames, info


def bind_host_to_gpu(local_index=0):
    """Pin the calling thread (and the threads it creates later) to the CPUs NVML reports as local to the GPU, so
    that pinned frame buffers are first-touched on the GPU's NUMA node: a D2H copy that has to cross the socket
    interconnect runs at ~38 GB/s instead of ~57 GB/s, i.e. costs a third of the end-to-end frame rate.  Call it
    before allocating pinned memory (ideally before importing torch).  Returns the number of CPUs bound to, or None
    when NVML is not available (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        index = int(local_index)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if visible:
            entries = [v.strip() for v in visible.split(",") if v.strip()]
            entry = entries[index]
            handle = (pynvml.nvmlDeviceGetHandleByUUID(entry.encode() if hasattr(entry, "encode") else entry)
                      if entry.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(entry)))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def window_create(nbytes):
    """Device buffer on this GPU that peer processes can map -> (device pointer, 64-byte IPC handle)."""
    lib = init()
    ptr, handle = C.c_void_p(), C.create_string_buffer(64)
    _check(lib.b2r_window_create(int(nbytes), C.byref(ptr), handle))
    return int(ptr.value), handle.raw


def window_open(handle: bytes):
    """Map a peer's window (its IPC handle) into this process -> device pointer."""
    lib = init()
    ptr = C.c_void_p()
    _check(lib.b2r_window_open(C.create_string_buffer(handle, 64), C.byref(ptr)))
    return int(ptr.value)


def window_push(src_ptr, dst_ptr, n_views, height, width, state_ptr, stream=0):
    """Sparse tile push of finished frames into a window block on `stream` (a cudaStream_t as int); see include/b2r.h."""
    _check(init().b2r_window_push(C.c_void_p(int(src_ptr)), C.c_void_p(int(dst_ptr)), int(n_views), int(height), int(width),
                                  C.c_void_p(int(state_ptr)), C.c_void_p(int(stream))))


def window_close(ptr):
    _check(init().b2r_window_close(C.c_void_p(ptr)))


def window_destroy(ptr):
    _check(init().b2r_window_destroy(C.c_void_p(ptr)))


class PendingFrames:
    """Frames of an asynchronous host render: `result()` blocks (GIL released) until they are in `out`."""

    def __init__(self, lib, ticket, frames, keep):
        self.lib, self.ticket, self.frames, self.keep = lib, ticket, frames, keep

    def result(self):
        if self.ticket is not None:
            _check(self.lib.b2r_wait(self.ticket))
            self.ticket = self.keep = None
        return self.frames


def status_report(models, face_status):
    """The three lines per model the reference prints during pass 3 (core.py:624-636)."""
    lines = []
    start = 0
    for m in models:
        n = len(m._faces)
        hist = np.bincount(np.asarray(face_status[start:start + n], dtype=np.uint8), minlength=32)
        start += n
        counts = {err: int(hist[err.value]) for err in Errors}
        lines += [f"Total faces {n}", f"Face rendered {int(hist[0])}", f"Discarded {counts}"]
    return lines
