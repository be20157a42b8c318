"""ctypes binding of the CPU oracle (oracle/libugsm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class Config(C.Structure):
    _fields_ = [
        ("levels", C.c_int32),
        ("fovea_level", C.c_int32),
        ("iters_cutoff", C.c_int32),
        ("smooth_passes", C.c_int32),
        ("smooth_passes_fine", C.c_int32),
        ("max_iters", C.c_int32),
    ]


def build():
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libugsm_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ugsmo_threshold_after.restype = C.c_float
        _LIB.ugsmo_threshold_after.argtypes = [C.c_int, C.c_int, C.c_float]
    return _LIB


def default_config(**kw):
    cfg = Config()
    lib().ugsmo_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def num_threads():
    return lib().ugsmo_num_threads()


def set_num_threads(n):
    lib().ugsmo_set_num_threads(int(n))


def use_all_host_cores():
    """OpenMP threads = the cores this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    set_num_threads(n)
    return num_threads()


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def taps():
    g = np.zeros(5, np.float32)
    a = np.zeros(5, np.float32)
    lib().ugsmo_taps(_ptr(g), _ptr(a))
    return g, a


def level_dims(cols, rows, levels=14):
    w = (C.c_int * 14)()
    h = (C.c_int * 14)()
    lib().ugsmo_level_dims(cols, rows, levels, w, h)
    return list(w)[:levels], list(h)[:levels]


def level_iters(cfg, i):
    return lib().ugsmo_level_iters(C.byref(cfg), i)


def level_smooth(cfg, i):
    return lib().ugsmo_level_smooth(C.byref(cfg), i)


def threshold_after(mi, m, T):
    return float(lib().ugsmo_threshold_after(mi, m, C.c_float(T)))


def unpack_rgb8(rgb):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    rows, cols, _ = rgb.shape
    out = np.empty((3, rows, cols), np.float32)
    lib().ugsmo_unpack_rgb8(_ptr(rgb), rows, cols, C.c_size_t(rgb.strides[0]), _ptr(out))
    return out


def _conv(fn, src, taps_):
    src = _f(src)
    h, w = src.shape
    dst = np.empty_like(src)
    t = _f(taps_)
    getattr(lib(), fn)(_ptr(src), _ptr(dst), w, h, _ptr(t))
    return dst


def conv_rows_zero(src, t):
    return _conv("ugsmo_conv_rows_zero", src, t)


def conv_cols_zero(src, t):
    return _conv("ugsmo_conv_cols_zero", src, t)


def conv_rows_clamp(src, t):
    return _conv("ugsmo_conv_rows_clamp", src, t)


def conv_cols_clamp(src, t):
    return _conv("ugsmo_conv_cols_clamp", src, t)


def decimate(src, scale, w2, h2):
    src = _f(src)
    h, w = src.shape
    dst = np.empty((h2, w2), np.float32)
    lib().ugsmo_decimate(_ptr(src), w, h, C.c_float(scale), _ptr(dst), w2, h2)
    return dst


def upsample_disp(src, w2, h2):
    src = _f(src)
    h, w = src.shape
    dst = np.empty((h2, w2), np.float32)
    lib().ugsmo_upsample_disp(_ptr(src), w, h, _ptr(dst), w2, h2)
    return dst


def warp(R, H, V):
    R, H, V = _f(R), _f(H), _f(V)
    h, w = R.shape
    dst = np.empty_like(R)
    sx = np.empty((h, w), np.int32)
    sy = np.empty((h, w), np.int32)
    lib().ugsmo_warp(_ptr(R), _ptr(H), _ptr(V), w, h, _ptr(dst), _ptr(sx), _ptr(sy))
    return dst, sx, sy


def poly(l, c, r, T):
    off = C.c_float()
    conf = C.c_float()
    br = C.c_uint8()
    lib().ugsmo_poly(C.c_float(l), C.c_float(c), C.c_float(r), C.c_float(T), C.byref(off), C.byref(conf), C.byref(br))
    return off.value, conf.value, br.value


def smooth_pass(src, conf):
    src, conf = _f(src), _f(conf)
    h, w = src.shape
    dst = src.copy()
    lib().ugsmo_smooth_pass(_ptr(src), _ptr(conf), _ptr(dst), w, h)
    return dst


def pyramid_plane(plane, levels=14):
    plane = _f(plane)
    rows, cols = plane.shape
    w, h = level_dims(cols, rows, levels)
    outs = [np.empty((h[l], w[l]), np.float32) for l in range(levels)]
    arr = (C.c_void_p * levels)(*[o.ctypes.data for o in outs])
    lib().ugsmo_pyramid_plane(_ptr(plane), cols, rows, levels, arr)
    return outs


def pyramid_rgb(rgb, levels=14):
    """-> list over levels of (3, h, w) float32 arrays."""
    planes = unpack_rgb8(rgb)
    per_ch = [pyramid_plane(planes[k], levels) for k in range(3)]
    return [np.stack([per_ch[k][l] for k in range(3)]) for l in range(levels)]


def correlate(L3, R3, H, V, Cc, T, first_coarsest=False, debug=False):
    L3, R3, H, V, Cc = _f(L3), _f(R3), _f(H), _f(V), _f(Cc)
    _, h, w = L3.shape
    Hn, Vn, Cn = np.empty_like(H), np.empty_like(V), np.empty_like(Cc)
    if debug:
        sx = np.empty((h, w), np.int32)
        sy = np.empty((h, w), np.int32)
        bh = np.empty((h, w), np.uint8)
        bv = np.empty((h, w), np.uint8)
        dbg = (_ptr(sx), _ptr(sy), _ptr(bh), _ptr(bv))
    else:
        dbg = (None, None, None, None)
    lib().ugsmo_correlate(_ptr(L3), _ptr(R3), _ptr(H), _ptr(V), _ptr(Cc), w, h, C.c_float(T), int(first_coarsest),
                          _ptr(Hn), _ptr(Vn), _ptr(Cn), *dbg)
    if debug:
        return Hn, Vn, Cn, sx, sy, bh, bv
    return Hn, Vn, Cn


def smooth(H, V, Cc, passes):
    H, V, Cc = _f(H).copy(), _f(V).copy(), _f(Cc).copy()
    h, w = H.shape
    lib().ugsmo_smooth(_ptr(H), _ptr(V), _ptr(Cc), w, h, passes)
    return H, V, Cc


def match_level(L3, R3, H, V, Cc, iters, passes, coarsest):
    L3, R3 = _f(L3), _f(R3)
    H, V, Cc = _f(H).copy(), _f(V).copy(), _f(Cc).copy()
    _, h, w = L3.shape
    lib().ugsmo_match_level(_ptr(L3), _ptr(R3), _ptr(H), _ptr(V), _ptr(Cc), w, h, iters, passes, int(coarsest))
    return H, V, Cc


def upsample_state(H, V, Cc, fine_w, fine_h, crop_x=0, crop_y=0, out_w=None, out_h=None):
    H, V, Cc = _f(H), _f(V), _f(Cc)
    h, w = H.shape
    out_w = fine_w if out_w is None else out_w
    out_h = fine_h if out_h is None else out_h
    outs = [np.empty((out_h, out_w), np.float32) for _ in range(3)]
    lib().ugsmo_upsample_state(_ptr(H), _ptr(V), _ptr(Cc), w, h, fine_w, fine_h, crop_x, crop_y, out_w, out_h,
                               *[_ptr(o) for o in outs])
    return tuple(outs)


def match_full(rgbL, rgbR, cfg=None):
    cfg = cfg or default_config()
    rgbL = np.ascontiguousarray(rgbL, np.uint8)
    rgbR = np.ascontiguousarray(rgbR, np.uint8)
    rows, cols, _ = rgbL.shape
    H = np.empty((rows, cols), np.float32)
    V = np.empty_like(H)
    Cc = np.empty_like(H)
    rc = lib().ugsmo_match_full(C.byref(cfg), _ptr(rgbL), _ptr(rgbR), rows, cols, C.c_size_t(rgbL.strides[0]),
                                _ptr(H), _ptr(V), _ptr(Cc))
    if rc != 0:
        raise ValueError("ugsmo_match_full: invalid arguments")
    return H, V, Cc


def match_foveated(rgbL, rgbR, cfg=None):
    cfg = cfg or default_config()
    rgbL = np.ascontiguousarray(rgbL, np.uint8)
    rgbR = np.ascontiguousarray(rgbR, np.uint8)
    rows, cols, _ = rgbL.shape
    w, h = level_dims(cols, rows, cfg.levels)
    fw, fh = w[cfg.fovea_level - 1], h[cfg.fovea_level - 1]
    shape = (cfg.fovea_level * fh, fw)
    sH = np.empty(shape, np.float32)
    sV = np.empty(shape, np.float32)
    sC = np.empty(shape, np.float32)
    ofw, ofh = C.c_int(), C.c_int()
    rc = lib().ugsmo_match_foveated(C.byref(cfg), _ptr(rgbL), _ptr(rgbR), rows, cols, C.c_size_t(rgbL.strides[0]),
                                    _ptr(sH), _ptr(sV), _ptr(sC), C.byref(ofw), C.byref(ofh))
    if rc != 0:
        raise ValueError("ugsmo_match_foveated: invalid arguments")
    assert (ofw.value, ofh.value) == (fw, fh)
    return sH, sV, sC, fw, fh


def partsubsample_disp(src, scale, w2, h2):
    src = _f(src)
    h, w = src.shape
    dst = np.empty((h2, w2), np.float32)
    lib().ugsmo_partsubsample_disp(_ptr(src), w, h, C.c_float(scale), _ptr(dst), w2, h2)
    return dst


def composite(sH, sV, sC, cols, rows, levels=14, fovea_level=7):
    sH, sV, sC = _f(sH), _f(sV), _f(sC)
    outs = [np.empty((rows, cols), np.float32) for _ in range(3)]
    lib().ugsmo_composite(_ptr(sH), _ptr(sV), _ptr(sC), cols, rows, levels, fovea_level, *[_ptr(o) for o in outs])
    return tuple(outs)


def match_composite(rgbL, rgbR, cfg=None):
    cfg = cfg or default_config()
    rgbL = np.ascontiguousarray(rgbL, np.uint8)
    rgbR = np.ascontiguousarray(rgbR, np.uint8)
    rows, cols, _ = rgbL.shape
    outs = [np.empty((rows, cols), np.float32) for _ in range(3)]
    rc = lib().ugsmo_match_composite(C.byref(cfg), _ptr(rgbL), _ptr(rgbR), rows, cols, C.c_size_t(rgbL.strides[0]),
                                     *[_ptr(o) for o in outs])
    if rc != 0:
        raise ValueError("ugsmo_match_composite: invalid arguments")
    return tuple(outs)


def poly_n(l, c, r, T):
    l, c, r = _f(l), _f(c), _f(r)
    n = l.size
    off, conf, br = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.uint8)
    lib().ugsmo_poly_n(_ptr(l), _ptr(c), _ptr(r), n, C.c_float(T), _ptr(off), _ptr(conf), _ptr(br))
    return off, conf, br


# ---- point cloud (oracle/ugsm_oracle_cloud.c) ----
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _cam(P):
    P = np.ascontiguousarray(P, np.float64).reshape(12)
    return P


def triangulate(P1, P2, x1, y1, x2, y2):
    out = np.empty(3, np.float32)
    lib().ugsmo_triangulate(_ptr(_cam(P1)), _ptr(_cam(P2)), C.c_float(x1), C.c_float(y1), C.c_float(x2),
                            C.c_float(y2), _ptr(out))
    return out


def cloud_count(rows, cols, sampling):
    f = lib().ugsmo_cloud_count
    f.restype = C.c_int64
    return int(f(rows, cols, sampling))


def cloud_margins(im_rows, im_cols, fovea_level, src_level):
    l, u = C.c_int(), C.c_int()
    lib().ugsmo_cloud_margins(im_rows, im_cols, fovea_level, src_level, C.byref(l), C.byref(u))
    return l.value, u.value


def cloud_scale(src_level):
    f = lib().ugsmo_cloud_scale
    f.restype = C.c_float
    return float(f(src_level))


def cloud_full(P1, P2, dX, dY, img, bgr=True, sampling=1, stride=4):
    """-> (n, stride) float32 records; column 3 (stride 4) / 4 (stride 8) holds the packed rgb bits."""
    dX, dY = _f(dX), _f(dY)
    img = np.ascontiguousarray(img, np.uint8)
    rows, cols = dX.shape
    n = cloud_count(rows, cols, sampling)
    pts = np.empty((n, stride), np.float32)
    f = lib().ugsmo_cloud_full
    f.restype = C.c_int64
    got = f(_ptr(_cam(P1)), _ptr(_cam(P2)), _ptr(dX), _ptr(dY), rows, cols, _ptr(img), C.c_size_t(img.strides[0]),
            int(bgr), sampling, _ptr(pts), stride)
    if got != n:
        raise ValueError("ugsmo_cloud_full: invalid arguments")
    return pts


def cloud_foveated(P1, P2, stackX, stackY, fov_h, src_level, im_rows, im_cols, img, fovea_level=7, bgr=True,
                   sampling=1, stride=4):
    stackX, stackY = _f(stackX), _f(stackY)
    img = np.ascontiguousarray(img, np.uint8)
    fov_w = stackX.shape[1]
    n = cloud_count(fov_h, fov_w, sampling)
    pts = np.empty((n, stride), np.float32)
    f = lib().ugsmo_cloud_foveated
    f.restype = C.c_int64
    got = f(_ptr(_cam(P1)), _ptr(_cam(P2)), _ptr(stackX), _ptr(stackY), fov_h, fov_w, fovea_level, src_level,
            im_rows, im_cols, _ptr(img), C.c_size_t(img.strides[0]), int(bgr), sampling, _ptr(pts), stride)
    if got != n:
        raise ValueError("ugsmo_cloud_foveated: invalid arguments")
    return pts


def range_full(P1, P2, dX, dY):
    """getRangePoint at every pixel (getPointCloud.cpp:951-982) -> (rows, cols) float32."""
    dX, dY = _f(dX), _f(dY)
    rows, cols = dX.shape
    out = np.empty((rows, cols), np.float32)
    if lib().ugsmo_range_full(_ptr(_cam(P1)), _ptr(_cam(P2)), _ptr(dX), _ptr(dY), rows, cols, _ptr(out)) != 0:
        raise ValueError("ugsmo_range_full: invalid arguments")
    return out


def range_foveated(P1, P2, stackX, stackY, fov_h, src_level, im_rows, im_cols, fovea_level=7):
    """getRangePointFOV at every fovea pixel of stack level src_level (getPointCloud.cpp:984-1019)."""
    stackX, stackY = _f(stackX), _f(stackY)
    fov_w = stackX.shape[1]
    out = np.empty((fov_h, fov_w), np.float32)
    if lib().ugsmo_range_foveated(_ptr(_cam(P1)), _ptr(_cam(P2)), _ptr(stackX), _ptr(stackY), fov_h, fov_w,
                                  fovea_level, src_level, im_rows, im_cols, _ptr(out)) != 0:
        raise ValueError("ugsmo_range_foveated: invalid arguments")
    return out


def run_refpc(mode, P1, P2, dX, dY, img, sampling=1, src_level=0, workdir=None):
    """Run oracle/_ref/refpc (the unmodified reference point-cloud source behind stubs).
    mode 0: full cloud; mode 1: foveated cloud of stack level src_level.  -> (n, 4) float32.
    mode 2 / 3: the range maps of the same inputs (getRangePoint / getRangePointFOV) -> (n,) float32, row-major."""
    import struct
    import tempfile

    exe = os.path.join(_HERE, "_ref", "refpc")
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    dX, dY = _f(dX), _f(dY)
    img = np.ascontiguousarray(img, np.uint8)
    with tempfile.TemporaryDirectory(dir=workdir) as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("7i", mode, dX.shape[0], dX.shape[1], img.shape[0], img.shape[1], sampling, src_level))
            f.write(_cam(P1).tobytes())
            f.write(_cam(P2).tobytes())
            f.write(dX.tobytes())
            f.write(dY.tobytes())
            f.write(img.tobytes())
        subprocess.run([exe, fin, fout], check=True)
        raw = open(fout, "rb").read()
    n = struct.unpack("q", raw[:8])[0]
    if mode >= 2:
        return np.frombuffer(raw[8:], np.float32, n).copy()
    return np.frombuffer(raw[8:], np.float32).reshape(n, 4).copy()
