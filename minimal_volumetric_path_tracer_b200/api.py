"""ctypes mirror of include/vpt.h.  Same names, argument meaning and error behaviour as the C boundary; numpy arrays stand
in for the caller-owned buffers.  Fails loudly (VptError / OSError) when the CUDA library is missing: there is no fallback."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvpt_b200.so")

METHOD_FREE_FLIGHT, METHOD_EQUIANGULAR, METHOD_MIS, METHOD_RAYMARCH, METHOD_MIS_DISTANCE, METHOD_VOLUME_SPHERES = 0, 1, 2, 3, 4, 5
PRECISION_FP32, PRECISION_FP64_REF = 0, 1
OUTPUT_SUM, OUTPUT_MEAN = 0, 1
KERNEL_AUTO, KERNEL_MEGA, KERNEL_WAVEFRONT, KERNEL_MEGA_SCAN, KERNEL_WAVEFRONT_SM, KERNEL_WAVEFRONT_HBM = 0, 1, 2, 3, 4, 5
QUIRK_R0_FALLTHROUGH, QUIRK_EXACT_VISIBILITY, QUIRKS_REFERENCE, QUIRKS_NONE = 1, 2, 3, 0


class UNIT:
    SPHERE_INTERSECT, INTERSECT, VISIBILITY, TRANSMITTANCE, FREE_FLIGHT, PHASE_SAMPLE, EQUIANGULAR, POWER_HEURISTIC = range(8)
    COSINE_HEMISPHERE, CONE_SAMPLE, MICROFACET, FACET_NORMAL, MEDIUM_NEE, POINT_LIGHT, SURFACE_MIS, BSDF_SAMPLE, RADIANCE, CAMERA_RAY, RADIANCE_LIST, RAYMARCH, MIS_DISTANCE, DIELECTRIC = range(8, 22)


class Sphere(C.Structure):  # vpt_sphere
    _fields_ = [("r", C.c_double), ("p", C.c_double * 3), ("c", C.c_double * 3), ("radiance", C.c_double * 3), ("material", C.c_int32),
                ("_pad", C.c_int32), ("eta", C.c_double * 3), ("kappa", C.c_double * 3), ("alpha", C.c_double)]


class Params(C.Structure):  # vpt_params
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("tile_rank", C.c_int32), ("tile_count", C.c_int32), ("method", C.c_int32), ("max_depth", C.c_int32),
                ("sigma_a", C.c_double), ("sigma_s", C.c_double), ("continue_prob", C.c_double),
                ("cam_o", C.c_double * 3), ("cam_dir", C.c_double * 3), ("fov", C.c_double), ("seed", C.c_uint64),
                ("quirks", C.c_uint32), ("precision", C.c_int32), ("output", C.c_int32), ("kernel", C.c_int32), ("device", C.c_int32),
                ("march_source", C.c_int32), ("march_step", C.c_double)]

    def copy(self, **kw):
        q = Params.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            if not hasattr(q, k):
                raise AttributeError(k)
            setattr(q, k, v)
        return q


class Stats(C.Structure):  # vpt_stats
    _fields_ = [("paths", C.c_uint64), ("events", C.c_uint64), ("scene_scans", C.c_uint64), ("nonfinite", C.c_uint64),
                ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VptError(RuntimeError):
    def __init__(self, status, text, detail=""):
        super().__init__("vpt error %d: %s%s" % (status, text, (" [" + detail + "]") if detail else ""))
        self.status = status


_lib = None


def load_library(path=None):
    """dlopen libvpt_b200.so (built by build.py).  Raises OSError if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise OSError("%s not found: build it first (python -m minimal_volumetric_path_tracer_b200.build); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    PP, PS, PF, PST = C.POINTER(Params), C.POINTER(Sphere), C.POINTER(C.c_float), C.POINTER(Stats)
    PD, PU32 = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    lib.vpt_default_params.argtypes = [PP]; lib.vpt_default_params.restype = None
    lib.vpt_default_scene.argtypes = [PS, C.c_int32]
    lib.vpt_load_scene.argtypes = [C.c_char_p, PS, C.c_int32]
    lib.vpt_render.argtypes = [PP, PS, C.c_int32, PF, PST]
    lib.vpt_render_device.argtypes = [PP, PS, C.c_int32, C.c_void_p, C.c_void_p, PST]
    lib.vpt_render_multi.argtypes = [PP, PS, C.c_int32, C.POINTER(C.c_int32), C.c_int32, PF, PST]
    lib.vpt_tonemap.argtypes = [PF, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
    lib.vpt_write_ppm.argtypes = [PF, C.c_int32, C.c_int32, C.c_char_p]
    lib.vpt_write_pfm.argtypes = [PF, C.c_int32, C.c_int32, C.c_char_p]
    lib.vpt_unit.argtypes = [C.c_int32, PP, PS, C.c_int32, C.c_int32, PD, C.c_int32, PD, C.c_int32]
    lib.vpt_unit_strides.argtypes = [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.vpt_philox.argtypes = [C.c_int32, C.c_int32, PU32, PU32, PU32]
    lib.vpt_measure_fp32_peak.argtypes = [C.c_int32, PD, PD]
    lib.vpt_strerror.restype = C.c_char_p; lib.vpt_strerror.argtypes = [C.c_int]
    lib.vpt_last_cuda_error.restype = C.c_char_p
    lib.vpt_version.restype = C.c_char_p
    lib.vpt_host_alloc.restype = C.c_void_p; lib.vpt_host_alloc.argtypes = [C.c_size_t]
    lib.vpt_host_free.restype = None; lib.vpt_host_free.argtypes = [C.c_void_p]
    if path == LIB_PATH:
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise VptError(rc, lib.vpt_strerror(rc).decode(), lib.vpt_last_cuda_error().decode() if rc in (-4, -5) else "")


def default_params(**kw):
    p = Params()
    load_library().vpt_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        if k in ("cam_o", "cam_dir"):
            v = (C.c_double * 3)(*v)
        setattr(p, k, v)
    return p


def default_scene():
    arr = (Sphere * 10)()
    n = load_library().vpt_default_scene(arr, 10)
    assert n == 10
    return arr


def load_scene(path):
    """a scene file (scenes/*.txt; format: include/vpt.h vpt_load_scene) -> Sphere array"""
    arr = (Sphere * 32)()
    lib = load_library()
    n = lib.vpt_load_scene(os.fsencode(path), arr, 32)
    if n < 0:
        _check(lib, n)
    return (Sphere * n).from_buffer_copy(bytes(arr)[: n * C.sizeof(Sphere)])


def scene_from_rows(rows):
    """rows: n x 18 (r, p3, c3, radiance3, material, eta3, kappa3, alpha) -- the reference's Sphere constructor order."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 18)
    arr = (Sphere * len(rows))()
    for i, d in enumerate(rows):
        s = arr[i]
        s.r = d[0]
        for k in range(3):
            s.p[k], s.c[k], s.radiance[k], s.eta[k], s.kappa[k] = d[1 + k], d[4 + k], d[7 + k], d[11 + k], d[14 + k]
        s.material = int(d[10]); s.alpha = d[17]
    return arr


def scene_to_rows(scene):
    out = np.zeros((len(scene), 18))
    for i, s in enumerate(scene):
        out[i] = [s.r, *s.p, *s.c, *s.radiance, s.material, *s.eta, *s.kappa, s.alpha]
    return out


def render(params, scene=None, stats=False):
    """vpt_render: host buffer out.  Returns hdr (h, w, 3) float32 [, Stats]."""
    lib = load_library()
    scene = scene if scene is not None else default_scene()
    hdr = np.empty((params.height, params.width, 3), dtype=np.float32)
    st = Stats()
    _check(lib, lib.vpt_render(C.byref(params), scene, len(scene), hdr.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
    return (hdr, st) if stats else hdr


def render_device(params, scene, hdr_ptr, stream=0, stats=None):
    """vpt_render_device: hdr_ptr is a device address (e.g. torch_tensor.data_ptr()); stream a cudaStream_t handle (int)."""
    lib = load_library()
    _check(lib, lib.vpt_render_device(C.byref(params), scene, len(scene), C.c_void_p(hdr_ptr), C.c_void_p(stream), C.byref(stats) if stats is not None else None))


def render_multi(params, scene, devices, stats=False):
    lib = load_library()
    scene = scene if scene is not None else default_scene()
    hdr = np.empty((params.height, params.width, 3), dtype=np.float32)
    dev = (C.c_int32 * len(devices))(*devices)
    st = Stats()
    _check(lib, lib.vpt_render_multi(C.byref(params), scene, len(scene), dev, len(devices), hdr.ctypes.data_as(C.POINTER(C.c_float)), C.byref(st)))
    return (hdr, st) if stats else hdr


def unit_strides(fn):
    a, b = C.c_int32(), C.c_int32()
    lib = load_library()
    _check(lib, lib.vpt_unit_strides(fn, C.byref(a), C.byref(b)))
    return a.value, b.value


def unit(fn, rows, params=None, scene=None):
    """vpt_unit: rows (n, in_stride) float64 -> (n, out_stride) float64."""
    lib = load_library()
    params = params if params is not None else default_params()
    scene = scene if scene is not None else default_scene()
    si, so = unit_strides(fn)
    rows = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1, si))
    out = np.zeros((len(rows), so), dtype=np.float64)
    PD = C.POINTER(C.c_double)
    _check(lib, lib.vpt_unit(fn, C.byref(params), scene, len(scene), len(rows), rows.ctypes.data_as(PD), si, out.ctypes.data_as(PD), so))
    return out


def philox(ctr, key, device=0):
    lib = load_library()
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4); key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
    out = np.zeros_like(ctr)
    P = C.POINTER(C.c_uint32)
    _check(lib, lib.vpt_philox(device, len(ctr), ctr.ctypes.data_as(P), key.ctypes.data_as(P), out.ctypes.data_as(P)))
    return out


def measure_fp32_peak(device=0):
    lib = load_library()
    t, c = C.c_double(), C.c_double()
    _check(lib, lib.vpt_measure_fp32_peak(device, C.byref(t), C.byref(c)))
    return t.value, c.value


def tonemap(hdr_mean):
    """mathUtilities.h:34-45: clamp to [0,1], gamma 2.2, *255 + .5, truncate.  Runs in the library's host code (not on the GPU)."""
    lib = load_library()
    hdr = np.ascontiguousarray(hdr_mean, dtype=np.float32)
    h, w = hdr.shape[:2]
    out = np.zeros((h, w, 3), dtype=np.uint8)
    _check(lib, lib.vpt_tonemap(hdr.ctypes.data_as(C.POINTER(C.c_float)), w, h, out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out


def write_ppm(hdr_mean, path):
    """rt.cpp:812-820: text P3, "%d %d %d " per pixel, no newlines after the header."""
    lib = load_library()
    hdr = np.ascontiguousarray(hdr_mean, dtype=np.float32)
    h, w = hdr.shape[:2]
    _check(lib, lib.vpt_write_ppm(hdr.ctypes.data_as(C.POINTER(C.c_float)), w, h, os.fsencode(path)))


def write_pfm(hdr_mean, path):
    """The frame before the tonemap as a binary colour PFM (little-endian floats, bottom row first): vpt_write_pfm."""
    lib = load_library()
    hdr = np.ascontiguousarray(hdr_mean, dtype=np.float32)
    h, w = hdr.shape[:2]
    _check(lib, lib.vpt_write_pfm(hdr.ctypes.data_as(C.POINTER(C.c_float)), w, h, os.fsencode(path)))


def read_pfm(path):
    """(h, w, 3) float32 frame of a colour PFM, row 0 = top of the image (the layout of every frame in this package)"""
    with open(path, "rb") as f:
        if f.readline().strip() != b"PF":
            raise ValueError("not a colour PFM: %s" % path)
        w, h = (int(t) for t in f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(w * h * 12), dtype="<f4" if scale < 0 else ">f4")
    if data.size != w * h * 3:
        raise ValueError("truncated PFM: %s" % path)
    return np.ascontiguousarray(data.reshape(h, w, 3)[::-1].astype(np.float32))


class PinnedFrame:
    """(h, w, 3) float32 frame in page-locked, device-mapped host memory (vpt_host_alloc): vpt_render stores its pixels straight into it
    while it renders.  `.array` is a numpy view; the memory is released when the object is dropped (or by close())."""

    def __init__(self, height, width):
        lib = load_library()
        self._lib, self.nbytes = lib, height * width * 3 * 4
        self._ptr = lib.vpt_host_alloc(self.nbytes)
        if not self._ptr:
            raise VptError(-5, "vpt_host_alloc failed", lib.vpt_last_cuda_error().decode())
        self.array = np.ctypeslib.as_array((C.c_float * (height * width * 3)).from_address(self._ptr)).reshape(height, width, 3)

    def close(self):
        if self._ptr:
            self.array = None
            self._lib.vpt_host_free(self._ptr); self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_into(params, scene, frame, stats=None):
    """vpt_render into a caller-owned host frame (numpy float32 (h, w, 3), e.g. PinnedFrame.array)."""
    lib = load_library()
    scene = scene if scene is not None else default_scene()
    assert frame.dtype == np.float32 and frame.shape == (params.height, params.width, 3) and frame.flags.c_contiguous
    _check(lib, lib.vpt_render(C.byref(params), scene, len(scene), frame.ctypes.data_as(C.POINTER(C.c_float)), C.byref(stats) if stats is not None else None))
    return frame


def device_count():
    return load_library().vpt_device_count()


def version():
    return load_library().vpt_version().decode()
