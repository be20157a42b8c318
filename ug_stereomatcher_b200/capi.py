"""ctypes binding of the C-ABI in include/ugsm.h (ug_stereomatcher_b200/libugsm.so).

There is no CPU fallback: if the library is missing or no sm_100 device is present the
calls raise.  Nothing here imports the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UGSM_LIB") or os.path.join(_HERE, "libugsm.so")  # UGSM_LIB: tuning builds
MAX_LEVELS = 14
TIMED_LEVELS = 4

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM = 0, 1, 2, 3, 4

# Every symbol include/ugsm.h declares (checked by tests/test_abi.py against the header).
SYMBOLS = [
    "ugsm_config_default", "ugsm_create", "ugsm_destroy", "ugsm_last_error", "ugsm_set_stream", "ugsm_query_dims",
    "ugsm_match_full", "ugsm_match_foveated", "ugsm_match_full_device", "ugsm_match_foveated_device",
    "ugsm_match_composite", "ugsm_match_composite_device", "ugsm_stage_composite", "ugsm_stage_poly",
    "ugsm_get_fovea_images", "ugsm_get_pyramid_level", "ugsm_get_level_disparity", "ugsm_get_timing",
    "ugsm_stage_pyramid", "ugsm_stage_correlate", "ugsm_stage_smooth", "ugsm_stage_upsample",
    "ugsm_stage_match_level", "ugsm_bench_correlate", "ugsm_bench_smooth", "ugsm_bench_null", "ugsm_version", "ugsm_band_begin",
    "ugsm_band_level_info", "ugsm_band_iteration", "ugsm_band_upsample", "ugsm_band_emit", "ugsm_band_pack",
    "ugsm_band_unpack", "ugsm_cloud_count", "ugsm_cloud_full", "ugsm_cloud_full_device", "ugsm_cloud_foveated",
    "ugsm_cloud_from_match", "ugsm_cloud_last_ms", "ugsm_bench_pyramid", "ugsm_host_alloc", "ugsm_host_free",
    "ugsm_rowband_match", "ugsm_rowband_last_error", "ugsm_range_full", "ugsm_range_foveated", "ugsm_range_from_match",
    "ugsm_feeder_open", "ugsm_feeder_info", "ugsm_feeder_next", "ugsm_feeder_release", "ugsm_feeder_close",
    "ugsm_feeder_last_error", "ugsm_tiff_info", "ugsm_tiff_read_bgr8",
]
FEEDER_END = 100


class Config(C.Structure):
    _fields_ = [
        ("levels", C.c_int32),
        ("fovea_level", C.c_int32),
        ("device", C.c_int32),
        ("iters_cutoff", C.c_int32),
        ("smooth_passes", C.c_int32),
        ("smooth_passes_fine", C.c_int32),
        ("max_iters", C.c_int32),
        ("use_graph", C.c_int32),
    ]


class Dims(C.Structure):
    _fields_ = [
        ("levels", C.c_int32),
        ("width", C.c_int32 * MAX_LEVELS),
        ("height", C.c_int32 * MAX_LEVELS),
        ("match_width", C.c_int32 * MAX_LEVELS),
        ("match_height", C.c_int32 * MAX_LEVELS),
        ("iters", C.c_int32 * MAX_LEVELS),
        ("smooth", C.c_int32 * MAX_LEVELS),
        ("fov_width", C.c_int32),
        ("fov_height", C.c_int32),
        ("stack_rows", C.c_int32),
        ("pixel_iterations", C.c_int64),
    ]


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_float),
        ("pyramid_ms", C.c_float),
        ("levels_ms", C.c_float),
        ("h2d_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("launches", C.c_int32),
        ("level_ms", C.c_float * MAX_LEVELS),
        ("k1_ms", C.c_float * MAX_LEVELS),
        ("k2_ms", C.c_float * MAX_LEVELS),
    ]


class BandLevel(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("pitch", C.c_int32),
        ("iters", C.c_int32),
        ("smooth", C.c_int32),
        ("state", C.c_void_p * 3),
    ]


class CameraPair(C.Structure):
    """sensor_msgs/CameraInfo.P of the left and right camera, row-major 3x4 (getPointCloud.cpp:1158-1159)."""
    _fields_ = [("P1", C.c_double * 12), ("P2", C.c_double * 12)]

    @classmethod
    def of(cls, P1, P2):
        c = cls()
        c.P1[:] = [float(v) for v in np.asarray(P1, np.float64).reshape(12)]
        c.P2[:] = [float(v) for v in np.asarray(P2, np.float64).reshape(12)]
        return c


class FeederPair(C.Structure):
    _fields_ = [("left", C.c_void_p), ("right", C.c_void_p), ("rows", C.c_int32), ("cols", C.c_int32),
                ("step", C.c_size_t), ("index", C.c_int64), ("slot", C.c_int32)]


class FeederInfo(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("step", C.c_size_t), ("list_images", C.c_int64),
                ("pairs", C.c_int64), ("slots", C.c_int32), ("pinned", C.c_int32), ("decode_ms", C.c_double)]


class UgsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ugsm error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def build():
    """Compile libugsm.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)


def lib():
    """Load libugsm.so.  Raises if it has not been built -- there is no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise UgsmError(-1, "libugsm.so not built (run __graft_entry__.build() or make -C "
                                "ug_stereomatcher_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ugsm_last_error.restype = C.c_char_p
        L.ugsm_last_error.argtypes = [C.c_void_p]
        L.ugsm_version.restype = C.c_char_p
        L.ugsm_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.ugsm_destroy.argtypes = [C.c_void_p]
        L.ugsm_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.ugsm_query_dims.argtypes = [C.POINTER(Config), C.c_int, C.c_int, C.c_int, C.POINTER(Dims)]
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.ugsm_match_full.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp]
        L.ugsm_match_foveated.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp, C.POINTER(i), C.POINTER(i)]
        L.ugsm_match_full_device.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp]
        L.ugsm_match_composite.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp]
        L.ugsm_match_composite_device.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp]
        L.ugsm_stage_composite.argtypes = [vp, vp, vp, vp, i, i, vp, vp, vp]
        L.ugsm_stage_poly.argtypes = [vp, vp, vp, vp, i, C.c_float, vp, vp, vp]
        L.ugsm_match_foveated_device.argtypes = [vp, vp, vp, i, i, sz, vp, vp, vp]
        L.ugsm_get_fovea_images.argtypes = [vp, i, vp, vp]
        L.ugsm_get_pyramid_level.argtypes = [vp, i, vp, vp]
        L.ugsm_get_level_disparity.argtypes = [vp, i, vp, vp, vp]
        L.ugsm_get_timing.argtypes = [vp, C.POINTER(Timing)]
        L.ugsm_stage_pyramid.argtypes = [vp, vp, i, i, sz, C.POINTER(vp)]
        L.ugsm_stage_correlate.argtypes = [vp, vp, vp, vp, vp, vp, i, i, C.c_float, i, vp, vp, vp, vp, vp, vp, vp]
        L.ugsm_stage_smooth.argtypes = [vp, vp, vp, vp, i, i, i]
        L.ugsm_stage_upsample.argtypes = [vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp, vp, vp]
        L.ugsm_stage_match_level.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, i]
        L.ugsm_bench_correlate.argtypes = [vp, i, i, i, C.POINTER(C.c_float)]
        L.ugsm_bench_smooth.argtypes = [vp, i, i, i, i, C.POINTER(C.c_float)]
        L.ugsm_bench_null.argtypes = [vp, i, C.POINTER(C.c_float)]
        L.ugsm_bench_pyramid.argtypes = [vp, i, i, i, i, C.POINTER(C.c_float)]
        L.ugsm_host_alloc.argtypes = [sz, C.POINTER(vp)]
        L.ugsm_host_free.argtypes = [vp]
        L.ugsm_band_begin.argtypes = [vp, vp, vp, i, i, sz, i]
        L.ugsm_band_level_info.argtypes = [vp, i, C.POINTER(BandLevel)]
        L.ugsm_band_iteration.argtypes = [vp, i, i, i, i, i]
        L.ugsm_band_upsample.argtypes = [vp, i, i, i]
        L.ugsm_band_emit.argtypes = [vp, vp, vp, vp, i, i]
        L.ugsm_band_pack.argtypes = [vp, i, i, i, vp]
        L.ugsm_band_unpack.argtypes = [vp, i, i, i, vp]
        i64p, cam = C.POINTER(C.c_int64), C.POINTER(CameraPair)
        L.ugsm_cloud_count.argtypes = [i, i, i, i64p]
        L.ugsm_cloud_full.argtypes = [vp, cam, vp, vp, i, i, vp, sz, i, i, vp, i, i64p]
        L.ugsm_cloud_full_device.argtypes = [vp, cam, vp, vp, i, i, i, vp, sz, i, i, vp, i, i64p]
        L.ugsm_cloud_foveated.argtypes = [vp, cam, vp, vp, i, i, i, i, i, vp, sz, i, i, vp, i, i64p]
        L.ugsm_cloud_from_match.argtypes = [vp, cam, i, i, vp, i, i, i64p]
        L.ugsm_cloud_last_ms.argtypes = [vp, C.POINTER(C.c_float)]
        ip = C.POINTER(C.c_int)
        L.ugsm_feeder_open.argtypes = [C.c_char_p, C.c_long, i, i, C.POINTER(vp)]
        L.ugsm_feeder_info.argtypes = [vp, C.POINTER(FeederInfo)]
        L.ugsm_feeder_next.argtypes = [vp, C.POINTER(FeederPair)]
        L.ugsm_feeder_release.argtypes = [vp, i]
        L.ugsm_feeder_close.argtypes = [vp]
        L.ugsm_feeder_last_error.argtypes = [vp]
        L.ugsm_feeder_last_error.restype = C.c_char_p
        L.ugsm_tiff_info.argtypes = [C.c_char_p, ip, ip, ip, ip]
        L.ugsm_tiff_read_bgr8.argtypes = [C.c_char_p, vp, sz, i, i]
        _LIB = L
    return _LIB


def default_config(**kw):
    cfg = Config()
    lib().ugsm_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def query_dims(cfg, rows, cols, mode=0):
    d = Dims()
    rc = lib().ugsm_query_dims(C.byref(cfg), rows, cols, mode, C.byref(d))
    if rc != OK:
        raise UgsmError(rc, "ugsm_query_dims: invalid rows/cols/mode/config")
    return d


def tiff_read(path):
    """ugsm_tiff_read_bgr8: the pixels of a TIFF file as cv::imread(path, 1) returns them (rows x cols x 3, B,G,R)."""
    L = lib()
    r, c, sp, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = L.ugsm_tiff_info(os.fsencode(path), C.byref(r), C.byref(c), C.byref(sp), C.byref(b))
    if rc != OK:
        raise UgsmError(rc, L.ugsm_feeder_last_error(None).decode())
    out = np.empty((r.value, c.value, 3), np.uint8)
    rc = L.ugsm_tiff_read_bgr8(os.fsencode(path), out.ctypes.data, c.value * 3, r.value, c.value)
    if rc != OK:
        raise UgsmError(rc, L.ugsm_feeder_last_error(None).decode())
    return out


class Feeder:
    """ugsm_feeder_*: image list -> TIFF decode threads -> ring of page-locked pair buffers (host-only)."""

    def __init__(self, list_xml, pairs=0, slots=4, threads=2):
        self._L = lib()
        self._f = C.c_void_p()
        rc = self._L.ugsm_feeder_open(os.fsencode(list_xml), pairs, slots, threads, C.byref(self._f))
        if rc != OK:
            self._f = None
            raise UgsmError(rc, self._L.ugsm_feeder_last_error(None).decode())

    def info(self):
        fi = FeederInfo()
        self._L.ugsm_feeder_info(self._f, C.byref(fi))
        return fi

    def next(self):
        """Returns (FeederPair, left, right) with numpy views of the slot (valid until release), or None at the end."""
        p = FeederPair()
        rc = self._L.ugsm_feeder_next(self._f, C.byref(p))
        if rc == FEEDER_END:
            return None
        if rc != OK:
            raise UgsmError(rc, self._L.ugsm_feeder_last_error(self._f).decode())
        n = p.rows * p.step
        left = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.left)).reshape(p.rows, p.cols, 3)
        right = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.right)).reshape(p.rows, p.cols, 3)
        return p, left, right

    def release(self, slot):
        rc = self._L.ugsm_feeder_release(self._f, slot)
        if rc != OK:
            raise UgsmError(rc, "ugsm_feeder_release: slot %d is not handed out" % slot)

    def close(self):
        if self._f:
            self._L.ugsm_feeder_close(self._f)
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def cloud_count(rows, cols, sampling=1):
    n = C.c_int64()
    if lib().ugsm_cloud_count(rows, cols, sampling, C.byref(n)) != OK:
        raise UgsmError(ERR_INVALID, "ugsm_cloud_count: invalid rows/cols/sampling")
    return n.value


def rowband_match(handles, devices, d_rgbL, d_rgbR, rows, cols, step, out, split_levels=3):
    """ugsm_rowband_match: one pair split into row bands over `handles` (one per GPU; device pointers as ints).
    `out`: three device pointers (dense rows x cols planes on devices[0]).  Returns the device time in ms."""
    n = len(handles)
    hs = (C.c_void_p * n)(*[h._h for h in handles])
    dv = (C.c_int * n)(*devices)
    pl = (C.c_void_p * n)(*d_rgbL)
    pr = (C.c_void_p * n)(*d_rgbR)
    ms = C.c_float()
    L = lib()
    L.ugsm_rowband_last_error.restype = C.c_char_p
    rc = L.ugsm_rowband_match(hs, dv, n, pl, pr, rows, cols, C.c_size_t(step), split_levels, C.c_void_p(out[0]),
                              C.c_void_p(out[1]), C.c_void_p(out[2]), C.byref(ms))
    if rc != OK:
        raise UgsmError(rc, L.ugsm_rowband_last_error().decode())
    return ms.value


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Handle:
    """One matcher instance: one CUDA stream, its own device buffers, not thread-safe."""

    def __init__(self, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self._h = C.c_void_p()
        rc = lib().ugsm_create(C.byref(self.cfg), C.byref(self._h))
        if rc != OK:
            raise UgsmError(rc, lib().ugsm_last_error(None).decode())

    def close(self):
        if self._h:
            lib().ugsm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != OK:
            raise UgsmError(rc, lib().ugsm_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        """cudaStream_t as an integer.  0 / None restores the handle's own stream (torch's default stream
        has handle 0: pass an explicit torch.cuda.Stream when ordering against other work matters)."""
        self._ck(lib().ugsm_set_stream(self._h, C.c_void_p(cuda_stream_ptr or None)))

    def timing(self):
        t = Timing()
        self._ck(lib().ugsm_get_timing(self._h, C.byref(t)))
        return t

    # -- whole-pair calls, host buffers ------------------------------------------------
    def match_full(self, rgbL, rgbR, out=None):
        rgbL = np.ascontiguousarray(rgbL, np.uint8)
        rgbR = np.ascontiguousarray(rgbR, np.uint8)
        rows, cols, ch = rgbL.shape
        if ch != 3 or rgbR.shape != rgbL.shape:
            raise UgsmError(ERR_INVALID, "expected two (rows, cols, 3) uint8 images of equal shape")
        if out is None:
            out = tuple(np.empty((rows, cols), np.float32) for _ in range(3))
        H, V, Cc = out
        for a in out:  # the library writes rows*cols floats through each pointer: a mismatch corrupts host memory
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == (rows, cols)
                                      and a.flags.c_contiguous and a.flags.writeable):
                raise UgsmError(ERR_INVALID, "out planes must be writable C-contiguous float32 arrays of shape "
                                             "(rows, cols) or None")
        self._ck(lib().ugsm_match_full(self._h, _p(rgbL), _p(rgbR), rows, cols, rgbL.strides[0], _p(H), _p(V), _p(Cc)))
        return H, V, Cc

    def match_foveated(self, rgbL, rgbR):
        rgbL = np.ascontiguousarray(rgbL, np.uint8)
        rgbR = np.ascontiguousarray(rgbR, np.uint8)
        rows, cols, ch = rgbL.shape
        if ch != 3 or rgbR.shape != rgbL.shape:
            raise UgsmError(ERR_INVALID, "expected two (rows, cols, 3) uint8 images of equal shape")
        d = query_dims(self.cfg, rows, cols, 1)
        shape = (d.stack_rows, d.fov_width)
        sH, sV, sC = (np.empty(shape, np.float32) for _ in range(3))
        fw, fh = C.c_int(), C.c_int()
        self._ck(lib().ugsm_match_foveated(self._h, _p(rgbL), _p(rgbR), rows, cols, rgbL.strides[0], _p(sH), _p(sV),
                                           _p(sC), C.byref(fw), C.byref(fh)))
        return sH, sV, sC, fw.value, fh.value

    def match_composite(self, rgbL, rgbR):
        """MatchGPULib::match(L, R, fov=1): foveated match + hierarchical composite -> full-resolution H, V, C."""
        rgbL = np.ascontiguousarray(rgbL, np.uint8)
        rgbR = np.ascontiguousarray(rgbR, np.uint8)
        rows, cols, ch = rgbL.shape
        if ch != 3 or rgbR.shape != rgbL.shape:
            raise UgsmError(ERR_INVALID, "expected two (rows, cols, 3) uint8 images of equal shape")
        H, V, Cc = (np.empty((rows, cols), np.float32) for _ in range(3))
        self._ck(lib().ugsm_match_composite(self._h, _p(rgbL), _p(rgbR), rows, cols, rgbL.strides[0], _p(H), _p(V),
                                            _p(Cc)))
        return H, V, Cc

    def stage_poly(self, l, c, r, T):
        l, c, r = map(_f32, (l, c, r))
        n = l.size
        off, conf, br = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
        self._ck(lib().ugsm_stage_poly(self._h, _p(l), _p(c), _p(r), n, T, _p(off), _p(conf), _p(br)))
        return off, conf, br

    def stage_composite(self, sH, sV, sC, rows, cols):
        sH, sV, sC = map(_f32, (sH, sV, sC))
        H, V, Cc = (np.empty((rows, cols), np.float32) for _ in range(3))
        self._ck(lib().ugsm_stage_composite(self._h, _p(sH), _p(sV), _p(sC), rows, cols, _p(H), _p(V), _p(Cc)))
        return H, V, Cc

    # -- point cloud from disparity planes (K5) -------------------------------------------
    def cloud_full(self, cam, H, V, img, bgr=True, sampling=1, stride=4):
        """doReconstructionRGB (getPointCloud.cpp:741-782) -> (n, stride) float32 records, column-major order."""
        H, V = _f32(H), _f32(V)
        img = np.ascontiguousarray(img, np.uint8)
        rows, cols = H.shape
        if V.shape != H.shape or img.shape != (rows, cols, 3):
            raise UgsmError(ERR_INVALID, "expected H, V of equal shape and a (rows, cols, 3) uint8 image")
        pts = np.empty((cloud_count(rows, cols, sampling), stride), np.float32)
        n = C.c_int64()
        self._ck(lib().ugsm_cloud_full(self._h, C.byref(cam), _p(H), _p(V), rows, cols, _p(img), img.strides[0],
                                       int(bgr), sampling, _p(pts), stride, C.byref(n)))
        assert n.value == pts.shape[0]
        return pts

    def cloud_foveated(self, cam, stackH, stackV, fov_rows, src_level, img, bgr=True, sampling=1, stride=4):
        """doReconstructionRGB_FOV (getPointCloud.cpp:697-740) for stack level src_level; img = full-res left image."""
        stackH, stackV = _f32(stackH), _f32(stackV)
        img = np.ascontiguousarray(img, np.uint8)
        fov_cols = stackH.shape[1]
        pts = np.empty((cloud_count(fov_rows, fov_cols, sampling), stride), np.float32)
        n = C.c_int64()
        self._ck(lib().ugsm_cloud_foveated(self._h, C.byref(cam), _p(stackH), _p(stackV), fov_rows, fov_cols, src_level,
                                           img.shape[0], img.shape[1], _p(img), img.strides[0], int(bgr), sampling,
                                           _p(pts), stride, C.byref(n)))
        assert n.value == pts.shape[0]
        return pts

    def range_full(self, cam, H, V):
        """getRangePoint at every pixel (getPointCloud.cpp:951-982) -> (rows, cols) float32."""
        H, V = _f32(H), _f32(V)
        rows, cols = H.shape
        out = np.empty((rows, cols), np.float32)
        self._ck(lib().ugsm_range_full(self._h, C.byref(cam), _p(H), _p(V), rows, cols, _p(out)))
        return out

    def range_foveated(self, cam, stackH, stackV, fov_rows, src_level, im_rows, im_cols):
        """getRangePointFOV at every fovea pixel of stack level src_level (getPointCloud.cpp:984-1019)."""
        stackH, stackV = _f32(stackH), _f32(stackV)
        fov_cols = stackH.shape[1]
        out = np.empty((fov_rows, fov_cols), np.float32)
        self._ck(lib().ugsm_range_foveated(self._h, C.byref(cam), _p(stackH), _p(stackV), fov_rows, fov_cols, src_level,
                                           im_rows, im_cols, _p(out)))
        return out

    def range_from_match(self, cam, rows, cols, src_level=-1):
        out = np.empty((rows, cols), np.float32)
        self._ck(lib().ugsm_range_from_match(self._h, C.byref(cam), src_level, _p(out), 0))
        return out

    def cloud_from_match(self, cam, rows, cols, src_level=-1, sampling=1, stride=4, out=None, d_points=None):
        """Cloud of the pair last matched on this handle, from the device-resident planes.  rows, cols: extent of
        the planes used (the image for src_level < 0, the fovea otherwise).  d_points: device pointer (int) to
        write to instead of returning a host array."""
        n = C.c_int64()
        if d_points is not None:
            self._ck(lib().ugsm_cloud_from_match(self._h, C.byref(cam), src_level, sampling, C.c_void_p(d_points),
                                                 stride, 1, C.byref(n)))
            return n.value
        pts = out if out is not None else np.empty((cloud_count(rows, cols, sampling), stride), np.float32)
        self._ck(lib().ugsm_cloud_from_match(self._h, C.byref(cam), src_level, sampling, _p(pts), stride, 0,
                                             C.byref(n)))
        assert n.value == pts.shape[0]
        return pts

    def cloud_full_device(self, cam, dH, dV, pitch, rows, cols, d_img, step, bgr, sampling, d_points, stride=4):
        n = C.c_int64()
        self._ck(lib().ugsm_cloud_full_device(self._h, C.byref(cam), dH, dV, pitch, rows, cols, d_img, step, int(bgr),
                                              sampling, d_points, stride, C.byref(n)))
        return n.value

    def cloud_last_ms(self):
        ms = C.c_float()
        self._ck(lib().ugsm_cloud_last_ms(self._h, C.byref(ms)))
        return ms.value

    # -- whole-pair calls, device pointers (ints) ----------------------------------------
    def match_full_device(self, d_rgbL, d_rgbR, rows, cols, step, dH, dV, dC):
        self._ck(lib().ugsm_match_full_device(self._h, d_rgbL, d_rgbR, rows, cols, step, dH, dV, dC))

    def match_foveated_device(self, d_rgbL, d_rgbR, rows, cols, step, dH, dV, dC):
        self._ck(lib().ugsm_match_foveated_device(self._h, d_rgbL, d_rgbR, rows, cols, step, dH, dV, dC))

    # -- inspection after a match ---------------------------------------------------------
    def pyramid_level(self, level, rows, cols, mode=0):
        d = query_dims(self.cfg, rows, cols, mode)
        shape = (3, d.height[level], d.width[level])
        L, R = np.empty(shape, np.float32), np.empty(shape, np.float32)
        self._ck(lib().ugsm_get_pyramid_level(self._h, level, _p(L), _p(R)))
        return L, R

    def fovea_images(self, level, rows, cols, mode=1):
        d = query_dims(self.cfg, rows, cols, mode)
        shape = (3, d.match_height[level], d.match_width[level])
        L, R = np.empty(shape, np.float32), np.empty(shape, np.float32)
        self._ck(lib().ugsm_get_fovea_images(self._h, level, _p(L), _p(R)))
        return L, R

    def level_disparity(self, level, rows, cols, mode=0):
        d = query_dims(self.cfg, rows, cols, mode)
        shape = (d.match_height[level], d.match_width[level])
        H, V, Cc = (np.empty(shape, np.float32) for _ in range(3))
        self._ck(lib().ugsm_get_level_disparity(self._h, level, _p(H), _p(V), _p(Cc)))
        return H, V, Cc

    # -- per-kernel stages ------------------------------------------------------------------
    def stage_pyramid(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        rows, cols, _ = rgb.shape
        d = query_dims(self.cfg, rows, cols, 0)
        outs = [np.empty((3, d.height[l], d.width[l]), np.float32) for l in range(d.levels)]
        arr = (C.c_void_p * d.levels)(*[o.ctypes.data for o in outs])
        self._ck(lib().ugsm_stage_pyramid(self._h, _p(rgb), rows, cols, rgb.strides[0], arr))
        return outs

    def stage_correlate(self, L3, R3, H, V, Cc, T, first_coarsest=False, debug=False):
        L3, R3, H, V, Cc = map(_f32, (L3, R3, H, V, Cc))
        _, h, w = L3.shape
        Hn, Vn, Cn = np.empty_like(H), np.empty_like(V), np.empty_like(Cc)
        if debug:
            sx, sy = np.empty((h, w), np.int32), np.empty((h, w), np.int32)
            bh, bv = np.empty((h, w), np.uint8), np.empty((h, w), np.uint8)
        else:
            sx = sy = bh = bv = None
        self._ck(lib().ugsm_stage_correlate(self._h, _p(L3), _p(R3), _p(H), _p(V), _p(Cc), w, h, T,
                                            int(first_coarsest), _p(Hn), _p(Vn), _p(Cn), _p(sx), _p(sy), _p(bh),
                                            _p(bv)))
        if debug:
            return Hn, Vn, Cn, sx, sy, bh, bv
        return Hn, Vn, Cn

    def stage_smooth(self, H, V, Cc, passes):
        H, V, Cc = _f32(H).copy(), _f32(V).copy(), _f32(Cc).copy()
        h, w = H.shape
        self._ck(lib().ugsm_stage_smooth(self._h, _p(H), _p(V), _p(Cc), w, h, passes))
        return H, V, Cc

    def stage_upsample(self, H, V, Cc, fine_w, fine_h, crop_x=0, crop_y=0, out_w=None, out_h=None):
        H, V, Cc = map(_f32, (H, V, Cc))
        h, w = H.shape
        out_w = fine_w if out_w is None else out_w
        out_h = fine_h if out_h is None else out_h
        outs = [np.empty((out_h, out_w), np.float32) for _ in range(3)]
        self._ck(lib().ugsm_stage_upsample(self._h, _p(H), _p(V), _p(Cc), w, h, fine_w, fine_h, crop_x, crop_y, out_w,
                                           out_h, *[_p(o) for o in outs]))
        return tuple(outs)

    def stage_match_level(self, L3, R3, H, V, Cc, iters, passes, coarsest):
        L3, R3 = _f32(L3), _f32(R3)
        H, V, Cc = _f32(H).copy(), _f32(V).copy(), _f32(Cc).copy()
        _, h, w = L3.shape
        self._ck(lib().ugsm_stage_match_level(self._h, _p(L3), _p(R3), _p(H), _p(V), _p(Cc), w, h, iters, passes,
                                              int(coarsest)))
        return H, V, Cc

    # -- row-band execution (device pointers; enqueue only, no synchronisation) -----------------
    def band_begin(self, d_rgbL, d_rgbR, rows, cols, step, split_levels):
        self._ck(lib().ugsm_band_begin(self._h, d_rgbL, d_rgbR, rows, cols, step, split_levels))

    def band_level_info(self, level):
        b = BandLevel()
        self._ck(lib().ugsm_band_level_info(self._h, level, C.byref(b)))
        return b

    def band_iteration(self, level, m, y_begin, y_end, ext):
        self._ck(lib().ugsm_band_iteration(self._h, level, m, y_begin, y_end, ext))

    def band_upsample(self, level, y_begin, y_end):
        self._ck(lib().ugsm_band_upsample(self._h, level, y_begin, y_end))

    def band_pack(self, level, y_begin, y_end, d_buf):
        self._ck(lib().ugsm_band_pack(self._h, level, y_begin, y_end, d_buf))

    def band_unpack(self, level, y_begin, y_end, d_buf):
        self._ck(lib().ugsm_band_unpack(self._h, level, y_begin, y_end, d_buf))

    def band_emit(self, dH, dV, dC, y_begin, y_end):
        self._ck(lib().ugsm_band_emit(self._h, dH, dV, dC, y_begin, y_end))

    def bench_correlate(self, width, height, reps=10):
        ms = C.c_float()
        self._ck(lib().ugsm_bench_correlate(self._h, width, height, reps, C.byref(ms)))
        return ms.value

    def bench_null(self, reps=100):
        ms = C.c_float()
        self._ck(lib().ugsm_bench_null(self._h, reps, C.byref(ms)))
        return ms.value

    def bench_pyramid(self, width, height, kind, reps=10):
        """kind 0: left energy (dense), 1: blur + odd-sample decimation, 2: blur + sqrt2 decimation."""
        ms = C.c_float()
        self._ck(lib().ugsm_bench_pyramid(self._h, width, height, kind, reps, C.byref(ms)))
        return ms.value

    def bench_smooth(self, width, height, passes, reps=10):
        ms = C.c_float()
        self._ck(lib().ugsm_bench_smooth(self._h, width, height, passes, reps, C.byref(ms)))
        return ms.value
