"""ctypes binding of the CPU oracle (oracle/fs_oracle.c) and of oracle/_ref (the reference's
vendored KissFFT compiled in place).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (audio-pathtracer_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_BANDS = 8


class Config(C.Structure):
    """fso_config (oracle/fs_oracle.h) -- layout-compatible with fs_config (include/frequensee.h)"""
    _fields_ = [
        ("n_bands", C.c_uint32), ("n_bins", C.c_uint32), ("bin_ms", C.c_float),
        ("rr_prob", C.c_float), ("eps_offset", C.c_float), ("eps_connect", C.c_float),
        ("min_seg", C.c_float), ("sound_speed", C.c_float), ("pdf_exponent", C.c_float),
        ("energy_clamp", C.c_float), ("energy_gain", C.c_float),
        ("air_absorption", C.c_float * MAX_BANDS),
        ("sample_rate", C.c_uint32), ("n_channels", C.c_uint32),
        ("ir_threshold", C.c_float), ("ir_lowpass", C.c_float),
        ("conv_block", C.c_uint32), ("conv_clamp", C.c_uint32), ("conv_wet", C.c_float),
        ("reserved", C.c_uint32 * 3),
    ]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("ext_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("connected", C.c_uint64), ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


PATH_DBG_DTYPE = np.dtype([
    ("n_src_nodes", np.uint32), ("n_lis_nodes", np.uint32), ("connected", np.uint32),
    ("bin", np.int32), ("delay_s", np.float32), ("total_dist", np.float32),
    ("energy", np.float32, (MAX_BANDS,)), ("src_end", np.float32, (3,)), ("lis_end", np.float32, (3,)),
])


def _has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def _build_if_needed():
    libs = [os.path.join(HERE, n) for n in ("liboracle_fma.so", "liboracle_generic.so")]
    src = os.path.join(HERE, "fs_oracle.c")
    stale = any((not os.path.exists(p)) or os.path.getmtime(p) < os.path.getmtime(src) for p in libs)
    if stale:
        subprocess.check_call(["make", "-C", HERE, "-s"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _build_if_needed()
        name = "liboracle_fma.so" if _has_fma() else "liboracle_generic.so"
        L = C.CDLL(os.path.join(HERE, name))
        fp = C.POINTER(C.c_float)
        L.fso_default_config.argtypes = [C.POINTER(Config)]
        L.fso_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        for n in ("fso_expf", "fso_logf", "fso_u01"):
            getattr(L, n).restype = C.c_float
        L.fso_expf.argtypes = [C.c_float]
        L.fso_logf.argtypes = [C.c_float]
        L.fso_u01.argtypes = [C.c_uint32]
        L.fso_powf.restype = C.c_float
        L.fso_powf.argtypes = [C.c_float, C.c_float]
        L.fso_sincos_2pi.argtypes = [C.c_float, fp, fp]
        L.fso_sample_sphere.argtypes = [C.c_float, C.c_float, fp]
        L.fso_sample_cos_hemisphere.argtypes = [fp, C.c_float, C.c_float, fp, fp]
        L.fso_intersect_tri.argtypes = [fp] * 6
        L.fso_intersect_tri.restype = C.c_int
        L.fso_scene_create.restype = C.c_void_p
        L.fso_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                       C.c_uint32, C.c_int]
        L.fso_scene_destroy.argtypes = [C.c_void_p]
        L.fso_closest_hit.argtypes = [C.c_void_p, fp, fp, fp, C.POINTER(C.c_uint32)]
        L.fso_closest_hit.restype = C.c_int
        L.fso_any_hit.argtypes = [C.c_void_p, fp, fp, C.c_float]
        L.fso_any_hit.restype = C.c_int
        L.fso_trace.restype = C.c_int
        L.fso_trace.argtypes = [C.c_void_p, C.POINTER(Config), C.c_void_p, C.c_uint32, C.c_void_p,
                                C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64,
                                C.c_void_p, C.POINTER(Stats), C.c_void_p, C.c_int]
        L.fso_scene_set_material_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fso_debug_path.argtypes = [C.c_void_p, C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32,
                                     C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
        L.fso_evaluate_nodes.argtypes = [C.c_void_p, C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, fp, C.c_void_p]
        L.fso_evaluate_nodes.restype = None
        L.fso_bin_index.argtypes = [C.POINTER(Config), C.c_float]
        L.fso_bin_index.restype = C.c_int32
        L.fso_build_ir.argtypes = [C.POINTER(Config), C.c_void_p, C.c_uint64, C.c_void_p]
        L.fso_build_ir_from_energy.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p]
        L.fso_band_carriers.argtypes = [C.POINTER(Config), C.c_uint64, C.c_void_p]
        L.fso_band_carriers.restype = None
        L.fso_build_ir_bands.argtypes = [C.POINTER(Config), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.fso_conv_create.restype = C.c_void_p
        L.fso_conv_create.argtypes = [C.POINTER(Config)]
        L.fso_conv_destroy.argtypes = [C.c_void_p]
        L.fso_conv_set_ir.argtypes = [C.c_void_p, C.c_void_p]
        L.fso_conv_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.fso_conv_process.restype = C.c_int
        _lib = L
    return _lib


def ref_lib():
    """oracle/_ref/libref_kissfft.so, or None when it was never built (reference tree absent)."""
    global _ref
    if _ref is None:
        p = os.path.join(HERE, "_ref", "libref_kissfft.so")
        if not os.path.exists(p):
            if os.path.isdir("/root/reference"):
                subprocess.check_call(["make", "-C", HERE, "-s", "ref"], stdout=subprocess.DEVNULL)
            if not os.path.exists(p):
                return None
        R = C.CDLL(p)
        R.ref_conv_create.restype = C.c_void_p
        R.ref_conv_create.argtypes = [C.c_int, C.c_int, C.c_int]
        R.ref_conv_destroy.argtypes = [C.c_void_p]
        R.ref_conv_fft_size.argtypes = [C.c_void_p]
        R.ref_conv_fft_size.restype = C.c_int
        R.ref_conv_set_ir.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_conv_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        R.ref_kiss_fftr.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        R.ref_kiss_fftri.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _ref = R
    return _ref


FLAG_CONNECT_ALL = 64
FLAG_SHARE_LISTENER = 128
FLAG_MATERIAL_MODEL = 256
FLAG_MIS = 512
FLAG_IR_NORMALIZE = 1024
FLAG_MIS_T1 = 1 << 20
FLAG_MIS_S1 = 1 << 21


def default_config(**over):
    cfg = Config()
    lib().fso_default_config(C.byref(cfg))
    for k, v in over.items():
        if k == "air_absorption":
            for i, a in enumerate(v):
                cfg.air_absorption[i] = a
        elif k == "flags":                                   # fs_config.flags lives in reserved[1]
            cfg.reserved[1] = int(v)
        else:
            setattr(cfg, k, v)
    return cfg


def _f3(a):
    return (C.c_float * 3)(*[float(x) for x in a])


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().fso_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return [int(x) for x in out]


def sincos_2pi(u):
    c, s = C.c_float(), C.c_float()
    lib().fso_sincos_2pi(u, C.byref(c), C.byref(s))
    return c.value, s.value


def sample_sphere(u1, u2):
    d = (C.c_float * 3)()
    lib().fso_sample_sphere(u1, u2, d)
    return np.array(d[:], dtype=np.float32)


def sample_cos_hemisphere(n, u1, u2):
    d = (C.c_float * 3)()
    ct = C.c_float()
    lib().fso_sample_cos_hemisphere(_f3(n), u1, u2, d, C.byref(ct))
    return np.array(d[:], dtype=np.float32), ct.value


class Scene:
    def __init__(self, verts, tri_mat, absorption, use_bvh=True):
        self.verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3, 3)
        self.tri_mat = np.ascontiguousarray(tri_mat, dtype=np.uint32)
        self.absorption = np.ascontiguousarray(absorption, dtype=np.float32)
        assert self.absorption.ndim == 2
        self.n_mats, self.n_bands = self.absorption.shape
        self.h = lib().fso_scene_create(self.verts.ctypes.data, self.tri_mat.ctypes.data,
                                        len(self.verts), self.absorption.ctypes.data,
                                        self.n_mats, self.n_bands, int(use_bvh))
        assert self.h

    def close(self):
        if self.h:
            lib().fso_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:                 # interpreter shutdown: the module globals may already be gone
            pass

    def set_material_model(self, transmission=None, scattering=None, thickness_cm=None):
        """transmission [M][B], scattering [M][B], thickness_cm [M] (None = 0 / 1 / 2.5 cm)"""
        def arr(a, shape):
            if a is None:
                return None
            a = np.ascontiguousarray(a, np.float32)
            assert a.shape == shape, (a.shape, shape)
            return a
        self._tr, self._sc, self._th = arr(transmission, self.absorption.shape), arr(scattering, self.absorption.shape), arr(thickness_cm, (self.n_mats,))
        p = lambda a: a.ctypes.data if a is not None else None      # noqa: E731
        lib().fso_scene_set_material_model(self.h, p(self._tr), p(self._sc), p(self._th))

    def debug_path(self, cfg, src, lis, g, n_paths, max_depth, seed):
        """node lists of path pair g: two arrays [n][8] = (p.xyz, ray origin.xyz, prob, bits) and the decoded (mat, event)"""
        s = np.ascontiguousarray(src, np.float32).reshape(3); l = np.ascontiguousarray(lis, np.float32).reshape(3)
        F = np.zeros((max_depth + 2, 8), np.float32); B = np.zeros((max_depth + 2, 8), np.float32)
        nf, nb = C.c_uint32(), C.c_uint32()
        lib().fso_debug_path(self.h, C.byref(cfg), s.ctypes.data, l.ctypes.data, g, n_paths, max_depth, seed,
                             F.ctypes.data, C.byref(nf), B.ctypes.data, C.byref(nb))
        out = []
        for a, n in ((F, nf.value), (B, nb.value)):
            bits = a[:n, 7].view(np.uint32)
            out.append({"p": a[:n, 0:3].copy(), "ro": a[:n, 3:6].copy(), "prob": a[:n, 6].copy(),
                        "mat": (bits & 0xffffff).astype(np.int64) - 1, "ev": (bits >> 24).astype(np.int64)})
        return out

    def closest_hit(self, o, d):
        t = C.c_float()
        tri = C.c_uint32()
        hit = lib().fso_closest_hit(self.h, _f3(o), _f3(d), C.byref(t), C.byref(tri))
        return (bool(hit), t.value, tri.value)

    def any_hit(self, o, d, tmax):
        return bool(lib().fso_any_hit(self.h, _f3(o), _f3(d), float(tmax)))

    def trace(self, cfg, src_pos, lis_pos, n_paths, max_depth, seed, g_first=0, g_count=None,
              n_threads=1, debug=False):
        src = np.ascontiguousarray(src_pos, dtype=np.float32).reshape(-1, 3)
        lis = np.ascontiguousarray(lis_pos, dtype=np.float32).reshape(3)
        S = len(src)
        if g_count is None:
            g_count = S * n_paths - g_first
        hist = np.zeros((S, cfg.n_bands, cfg.n_bins), dtype=np.uint64)
        st = Stats()
        dbg = np.zeros(g_count, dtype=PATH_DBG_DTYPE) if debug else None
        rc = lib().fso_trace(self.h, C.byref(cfg), src.ctypes.data, S, lis.ctypes.data, n_paths,
                             g_first, g_count, max_depth, seed, hist.ctypes.data, C.byref(st),
                             dbg.ctypes.data if debug else None, n_threads)
        if rc != 0:
            raise RuntimeError("fso_trace failed: %d" % rc)
        return (hist, st.as_dict(), dbg) if debug else (hist, st.as_dict())


def evaluate_nodes(scene, cfg, pos, mat, prob):
    """EvaluatePath (SUB.cpp:360-420) on an explicit node list -> (delay_s, energy[B])"""
    pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
    mat = np.ascontiguousarray(mat, dtype=np.int32)
    prob = np.ascontiguousarray(prob, dtype=np.float32)
    d = C.c_float()
    e = np.zeros(cfg.n_bands, np.float32)
    lib().fso_evaluate_nodes(scene.h, C.byref(cfg), pos.ctypes.data, mat.ctypes.data, prob.ctypes.data, len(pos), C.byref(d), e.ctypes.data)
    return d.value, e


def bin_index(cfg, delay_s):
    return int(lib().fso_bin_index(C.byref(cfg), float(delay_s)))


def build_ir(cfg, hist, n_paths):
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    out = np.zeros((cfg.n_channels, cfg.sample_rate), dtype=np.float32)
    lib().fso_build_ir(C.byref(cfg), hist.ctypes.data, n_paths, out.ctypes.data)
    return out


def band_carriers(cfg, seed):
    out = np.zeros((cfg.n_channels, cfg.n_bands, cfg.sample_rate), dtype=np.float32)
    lib().fso_band_carriers(C.byref(cfg), seed, out.ctypes.data)
    return out


def build_ir_bands(cfg, hist, n_paths, noise_seed):
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    out = np.zeros((cfg.n_channels, cfg.sample_rate), dtype=np.float32)
    lib().fso_build_ir_bands(C.byref(cfg), hist.ctypes.data, n_paths, noise_seed, out.ctypes.data)
    return out


def build_ir_from_energy(cfg, energy):
    energy = np.ascontiguousarray(energy, dtype=np.float32)
    out = np.zeros((cfg.n_channels, cfg.sample_rate), dtype=np.float32)
    lib().fso_build_ir_from_energy(C.byref(cfg), energy.ctypes.data, out.ctypes.data)
    return out


class Conv:
    """streaming direct-form double-precision convolver (semantics of REV.cpp:118-213)"""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().fso_conv_create(C.byref(cfg))

    def set_ir(self, ir):
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        assert ir.shape == (self.cfg.n_channels, self.cfg.sample_rate)
        lib().fso_conv_set_ir(self.h, ir.ctypes.data)

    def process(self, block):
        block = np.ascontiguousarray(block, dtype=np.float32)
        out = np.zeros_like(block)
        rc = lib().fso_conv_process(self.h, block.ctypes.data, out.ctypes.data, block.shape[0])
        assert rc == 0
        return out

    def __del__(self):
        if self.h:
            lib().fso_conv_destroy(self.h)
            self.h = None


class RefKissConv:
    """the reference's own convolution scheme on its own KissFFT (oracle/_ref)"""

    def __init__(self, sample_rate=48000, frame=1024, channels=2):
        self.R = ref_lib()
        if self.R is None:
            raise RuntimeError("oracle/_ref not built")
        self.frame, self.channels, self.sample_rate = frame, channels, sample_rate
        self.h = self.R.ref_conv_create(sample_rate, frame, channels)

    @property
    def fft_size(self):
        return self.R.ref_conv_fft_size(self.h)

    def set_ir(self, ir):
        ir = np.ascontiguousarray(ir, dtype=np.float32)
        self.R.ref_conv_set_ir(self.h, ir.ctypes.data)

    def process(self, block, clamp=True):
        block = np.ascontiguousarray(block, dtype=np.float32)
        out = np.zeros_like(block)
        self.R.ref_conv_process(self.h, block.ctypes.data, out.ctypes.data, int(clamp))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.R.ref_conv_destroy(self.h)
            self.h = None


def kiss_fftr(x):
    R = ref_lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros((len(x) // 2 + 1, 2), dtype=np.float32)
    R.ref_kiss_fftr(len(x), x.ctypes.data, out.ctypes.data)
    return out[:, 0] + 1j * out[:, 1]


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref/libref_ue_bodies.so: the REFERENCE's own UpdateSource / GenerateFullPaths / ConnectSubpaths / GeneratePath /
# EvaluatePath / AddEnergyAtDelay / ReconstructImpulseResponse / ConvolveFFT / FCircularAudioBuffer, compiled from
# /root/reference against oracle/ue_shim (see oracle/ref_ue_bodies.cpp).  Used to pin fs_oracle.c.
# ---------------------------------------------------------------------------------------------------------------------
_ue = None

UE_PATH_REC = np.dtype([("n_src_nodes", np.uint32), ("n_lis_nodes", np.uint32), ("connected", np.uint32), ("pad", np.uint32),
                        ("delay_s", np.float32), ("gain", np.float32), ("src_end", np.float64, (3,)), ("lis_end", np.float64, (3,))])


def ref_ue_lib():
    """None when oracle/_ref/libref_ue_bodies.so was never built (reference tree absent and no prebuilt copy)"""
    global _ue
    if _ue is None:
        p = os.path.join(HERE, "_ref", "libref_ue_bodies.so")
        if not os.path.exists(p):
            if os.path.isdir("/root/reference"):
                subprocess.check_call(["make", "-C", HERE, "-s", "ref_ue"], stdout=subprocess.DEVNULL)
            if not os.path.exists(p):
                return None
        U = C.CDLL(p)
        vp = C.c_void_p
        U.ref_ue_world_create.restype = vp
        U.ref_ue_world_create.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32, C.c_double]
        U.ref_ue_world_destroy.argtypes = [vp]
        U.ref_ue_world_destroy.restype = None
        U.ref_ue_update_source.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, vp, vp, C.POINTER(C.c_uint64)]
        U.ref_ue_paths.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.c_int, vp]
        U.ref_ue_evaluate_path.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        U.ref_ue_add_energy.argtypes = [vp, vp, C.c_int, vp]
        U.ref_ue_reconstruct_ir.argtypes = [vp, vp]
        U.ref_ue_conv_create.restype = vp
        U.ref_ue_conv_create.argtypes = [C.c_int, C.c_int]
        U.ref_ue_conv_destroy.argtypes = [vp]
        U.ref_ue_conv_destroy.restype = None
        U.ref_ue_conv_fft_size.argtypes = [vp]
        U.ref_ue_conv_set_ir.argtypes = [vp, vp]
        U.ref_ue_conv_set_ir.restype = None
        U.ref_ue_conv_process.argtypes = [vp, vp, vp, C.c_int]
        U.ref_ue_conv_process.restype = None
        _ue = U
    return _ue


# the reference's compile-time constants as an oracle configuration: one band, Absorption[2] used as
# reflectivity, distances in units of 1000 world units, "NodeDistance < 1" skip, air 0.05, offsets 0.1 world units
UE_UNIT = 1000.0          # world units per metre in the pin tests, so that EvaluatePath's "/ 1000" yields metres


def ue_pin_config(**over):
    kw = dict(n_bands=1, n_bins=1000, bin_ms=1.0, rr_prob=0.9, eps_offset=0.1 / UE_UNIT, eps_connect=0.1 / UE_UNIT, min_seg=1.0,
              sound_speed=343.0, pdf_exponent=0.1, energy_clamp=1.0, energy_gain=10.0, air_absorption=[0.05] * MAX_BANDS)
    kw.update(over)
    return default_config(**kw)


class RefUEWorld:
    """a triangle scene inside the engine stand-in; value2[m] = Absorption[2].Value of material m (used as reflectivity)"""

    def __init__(self, verts, tri_mat, value2, unit=UE_UNIT):
        self.U = ref_ue_lib()
        if self.U is None:
            raise RuntimeError("oracle/_ref/libref_ue_bodies.so not built")
        v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3, 3)
        m = np.ascontiguousarray(tri_mat, np.uint32)
        a = np.ascontiguousarray(value2, np.float32)
        self.unit = unit
        self.h = self.U.ref_ue_world_create(v.ctypes.data, m.ctypes.data, len(v), a.ctypes.data, len(a), unit)

    def update_source(self, src, lis, seed, g_first=0):
        """the reference's UpdateSource, whole and unmodified: (EnergyBuffer[1000], ImpulseBuffer[2][48000], line traces)"""
        s, l = np.ascontiguousarray(src, np.float32).reshape(3), np.ascontiguousarray(lis, np.float32).reshape(3)
        e = np.zeros(1000, np.float32); ir = np.zeros((2, 48000), np.float32); n = C.c_uint64()
        rc = self.U.ref_ue_update_source(self.h, s.ctypes.data, l.ctypes.data, seed, g_first, e.ctypes.data, ir.ctypes.data, C.byref(n))
        assert rc == 0
        return e, ir, n.value

    def paths(self, src, lis, seed, n, g_first=0):
        s, l = np.ascontiguousarray(src, np.float32).reshape(3), np.ascontiguousarray(lis, np.float32).reshape(3)
        rec = np.zeros(n, UE_PATH_REC)
        rc = self.U.ref_ue_paths(self.h, s.ctypes.data, l.ctypes.data, seed, g_first, n, rec.ctypes.data)
        assert rc == 0, rc
        return rec

    def close(self):
        if getattr(self, "h", None):
            self.U.ref_ue_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_ue_evaluate_path(pos_units, value2, prob):
    U = ref_ue_lib()
    p = np.ascontiguousarray(pos_units, np.float64).reshape(-1, 3)
    a = np.ascontiguousarray(value2, np.float32); q = np.ascontiguousarray(prob, np.float32)
    d, g = C.c_float(), C.c_float()
    U.ref_ue_evaluate_path(p.ctypes.data, a.ctypes.data, q.ctypes.data, len(p), C.byref(d), C.byref(g))
    return d.value, g.value


def ref_ue_add_energy(delays, energies):
    U = ref_ue_lib()
    d = np.ascontiguousarray(delays, np.float32); e = np.ascontiguousarray(energies, np.float32)
    out = np.zeros(1000, np.float32)
    U.ref_ue_add_energy(d.ctypes.data, e.ctypes.data, len(d), out.ctypes.data)
    return out


def ref_ue_reconstruct_ir(energy):
    """-> (ir[2][48000], the NumSamplesPerBin the reference computes)"""
    U = ref_ue_lib()
    e = np.ascontiguousarray(energy, np.float32)
    ir = np.zeros((2, 48000), np.float32)
    spb = U.ref_ue_reconstruct_ir(e.ctypes.data, ir.ctypes.data)
    return ir, spb


class RefUEConv:
    """the reference's reverb: its Initialize + ConvolveFFT bodies, its FCircularAudioBuffer and its KissFFT"""

    def __init__(self, sample_rate=48000, frame=1024):
        self.U = ref_ue_lib()
        if self.U is None:
            raise RuntimeError("oracle/_ref/libref_ue_bodies.so not built")
        self.h = self.U.ref_ue_conv_create(sample_rate, frame)

    @property
    def fft_size(self):
        return self.U.ref_ue_conv_fft_size(self.h)

    def set_ir(self, ir):
        ir = np.ascontiguousarray(ir, np.float32)
        self.U.ref_ue_conv_set_ir(self.h, ir.ctypes.data)

    def process(self, block, clamp=True):
        block = np.ascontiguousarray(block, np.float32)
        out = np.zeros_like(block)
        self.U.ref_ue_conv_process(self.h, block.ctypes.data, out.ctypes.data, int(clamp))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.U.ref_ue_conv_destroy(self.h)
            self.h = None
