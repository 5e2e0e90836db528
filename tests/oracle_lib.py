"""ctypes bindings for the two CPU checkers under oracle/ (TEST INFRASTRUCTURE).

L1 = oracle/_build/libvpt_oracle.so : our FP64 restatement (oracle/vpt_oracle.hpp), always buildable.
L0 = oracle/_ref/libvpt_l0.so       : the unmodified reference headers behind oracle/l0_harness.cpp,
                                       only buildable where /root/reference exists (prebuilt file travels).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
L1_PATH = os.path.join(ORACLE, "_build", "libvpt_oracle.so")
L0_PATH = os.path.join(ORACLE, "_ref", "libvpt_l0.so")
REF_RT_PATH = os.path.join(ORACLE, "_ref", "ref_rt")

D = C.c_double
PD = C.POINTER(D)
PU32 = C.POINTER(C.c_uint32)
PU64 = C.POINTER(C.c_uint64)

QUIRK_R0_FALLTHROUGH = 1
QUIRK_EXACT_VISIBILITY = 2
QUIRKS_REFERENCE = 3
QUIRKS_ROBUST = 0

# include/Sphere.cpp:11-22 restated as data: r, p, c, radiance, material, eta, kappa, alpha
_AL_ETA = (1.66058, 0.88143, 0.521467)
_AL_KAPPA = (9.2282, 6.27077, 4.83803)
DEFAULT_SCENE = np.array([
    [1e5, -1e5 - 49, 0, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1e5, 1e5 + 49, 0, 0, .0, .0, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1e5, 0, 0, -1e5 - 81.6, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1e5, 0, -1e5 - 40.8, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [1e5, 0, 1e5 + 40.8, 0, .5, .5, .5, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [16.5, -23, -24.3, -34.6, 0, 0, 0, 0, 0, 0, 1, *_AL_ETA, *_AL_KAPPA, 0.09],
    [16.5, 23, -24.3, -3.6, .0, .0, .9, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [2, 0, 24.3, -35, 0, 0, 0, 100, 100, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [0, -23, 24.3, 0, 0, 0, 0, 6000, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
    [2, 23, 24.3, 35, 0, 0, 0, 75, 75, 60, 0, 0, 0, 0, 0, 0, 0, 0],
], dtype=np.float64)
CAM_O = (0.0, 11.2, 214.0)            # src/rt.cpp:755
CAM_DIR = (0.0, -0.042612, -1.0)
CAM_FOV = 0.5095                      # src/rt.cpp:758


def scene_without(indices, scene=DEFAULT_SCENE):
    keep = [i for i in range(len(scene)) if i not in set(indices)]
    return np.ascontiguousarray(scene[keep])


def build_oracles(quiet=True):
    """(Re)build the checkers; L0 only where the reference tree exists (oracle/Makefile decides)."""
    out = subprocess.run(["make", "-C", ORACLE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _p(a):
    return a.ctypes.data_as(PD)


def _v(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class L1:
    """Our FP64 restatement.  Methods mirror the reference's function names."""

    def __init__(self, path=L1_PATH):
        if not os.path.exists(path):
            build_oracles()
        self.lib = C.CDLL(path)
        L = self.lib
        for name in ("l1_sphere_intersect", "l1_cosinethetaMax", "l1_transmitance", "l1_freeFlightSample", "l1_freeFlightProb",
                     "l1_pdfSuccess", "l1_pdfFailure", "l1_equiAngularSample", "l1_equiAngularProb", "l1_equiAngularParams2",
                     "l1_solidAngleProb", "l1_hemiCosineProb", "l1_NDF", "l1_G_smith", "l1_microFacetProb", "l1_fresnelDie",
                     "l1_powerHeuristics"):
            getattr(L, name).restype = D
        L.l1_radiance_erand48.restype = C.c_uint64

    # -- helpers
    @staticmethod
    def _u(u):
        u = _v(u)
        return u, _p(u), C.c_int(len(u))

    def philox(self, ctr, key):
        c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        self.lib.l1_philox(c.ctypes.data_as(PU32), k.ctypes.data_as(PU32), o.ctypes.data_as(PU32))
        return o

    def check_scene(self, scene):
        s = _v(scene); return self.lib.l1_check_scene(_p(s), C.c_int(len(s)))

    # -- geometry
    def sphere_intersect(self, sph, o, d):
        s, o, d = _v(sph), _v(o), _v(d); return self.lib.l1_sphere_intersect(_p(s), _p(o), _p(d))

    def intersect(self, scene, quirks, o, d, id0=0):
        s, o, d = _v(scene), _v(o), _v(d); t = D(0); i = C.c_int(id0)
        h = self.lib.l1_intersect(_p(s), C.c_int(len(s)), C.c_uint(quirks), _p(o), _p(d), C.byref(t), C.byref(i))
        return bool(h), t.value, i.value

    def visibility(self, scene, quirks, light, x):
        s, a, b = _v(scene), _v(light), _v(x)
        return bool(self.lib.l1_visibility(_p(s), C.c_int(len(s)), C.c_uint(quirks), _p(a), _p(b)))

    def cosinethetaMax(self, scene, i, x):
        s, x = _v(scene), _v(x); return self.lib.l1_cosinethetaMax(_p(s), C.c_int(len(s)), C.c_int(i), _p(x))

    def coordinateSystem(self, n):
        n = _v(n); s = np.zeros(3); t = np.zeros(3); self.lib.l1_coordinateSystem(_p(n), _p(s), _p(t)); return s, t

    def coordinateTraspose(self, n, w):
        n = _v(n); w = _v(w).copy(); self.lib.l1_coordinateTraspose(_p(n), _p(w)); return w

    # -- volume
    def transmitance(self, a, b, st):
        a, b = _v(a), _v(b); return self.lib.l1_transmitance(_p(a), _p(b), D(st))

    def freeFlightSample(self, st, u):
        u, pu, nu = self._u(u); return self.lib.l1_freeFlightSample(D(st), pu, nu)

    def freeFlightProb(self, st, d): return self.lib.l1_freeFlightProb(D(st), D(d))
    def pdfSuccess(self, st, t): return self.lib.l1_pdfSuccess(D(st), D(t))
    def pdfFailure(self, st, t): return self.lib.l1_pdfFailure(D(st), D(t))

    def isotropicPhaseSample(self, u):
        u, pu, nu = self._u(u); o = np.zeros(3); self.lib.l1_isotropicPhaseSample(pu, nu, _p(o)); return o

    def equiAngularSample(self, Dd, a, b, u):
        u, pu, nu = self._u(u); return self.lib.l1_equiAngularSample(D(Dd), D(a), D(b), pu, nu)

    def equiAngularProb(self, Dd, a, b, t): return self.lib.l1_equiAngularProb(D(Dd), D(a), D(b), D(t))

    def equiAngularParams2(self, scene, src, tmax, o, d, u):
        s, o, d = _v(scene), _v(o), _v(d); u, pu, nu = self._u(u); out = np.zeros(4)
        r = self.lib.l1_equiAngularParams2(_p(s), C.c_int(len(s)), C.c_int(src), D(tmax), _p(o), _p(d), pu, nu, _p(out))
        return r, out

    def dielectric(self, rows):
        """material 2's refraxDielectric / reflexDielectric / fresnelDie; rows: n x 6 (n[3], wo[3]) -> n x 7 (wt[3], wr[3], F)"""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 6); out = np.zeros((len(rows), 7))
        self.lib.l1_dielectric(C.c_int(len(rows)), _p(rows), _p(out)); return out

    def mis_distance(self, scene, rows):
        """method 4's distance decision; rows: n x 11 (source, tMax, o[3], d[3], sigma_t, xi, xd) -> n x 3 (surface, dist, mixture pdf)"""
        s = _v(scene); rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 11); out = np.zeros((len(rows), 3))
        self.lib.l1_mis_distance(_p(s), C.c_int(len(s)), C.c_int(len(rows)), _p(rows), _p(out)); return out

    def freeSingleScattering(self, scene, quirks, xt, src, st, pS, u):
        s, xt = _v(scene), _v(xt); u, pu, nu = self._u(u); o = np.zeros(3)
        self.lib.l1_freeSingleScattering(_p(s), C.c_int(len(s)), C.c_uint(quirks), _p(xt), C.c_int(src), D(st), D(pS), pu, nu, _p(o)); return o

    def singleScattering(self, scene, quirks, xt, src, st, ss, T, pS, u):
        s, xt = _v(scene), _v(xt); u, pu, nu = self._u(u); o = np.zeros(3)
        self.lib.l1_singleScattering(_p(s), C.c_int(len(s)), C.c_uint(quirks), _p(xt), C.c_int(src), D(st), D(ss), D(T), D(pS), pu, nu, _p(o)); return o

    # -- surface
    def cosineHemispheric(self, n, u):
        n = _v(n); u, pu, nu = self._u(u); o = np.zeros(3); self.lib.l1_cosineHemispheric(_p(n), pu, nu, _p(o)); return o

    def solidAngleDir(self, wc, cmax, u):
        wc = _v(wc); u, pu, nu = self._u(u); o = np.zeros(3); self.lib.l1_solidAngleDir(_p(wc), D(cmax), pu, nu, _p(o)); return o

    def solidAngleProb(self, c): return self.lib.l1_solidAngleProb(D(c))
    def hemiCosineProb(self, c): return self.lib.l1_hemiCosineProb(D(c))

    def vectorFacet(self, alpha, u):
        u, pu, nu = self._u(u); o = np.zeros(3); self.lib.l1_vectorFacet(D(alpha), pu, nu, _p(o)); return o

    def NDF(self, c, a): return self.lib.l1_NDF(D(c), D(a))

    def fresnel(self, c, eta, kappa):
        e, k = _v(eta), _v(kappa); o = np.zeros(3); self.lib.l1_fresnel(D(c), _p(e), _p(k), _p(o)); return o

    def G_smith(self, n, wi, wo, wh, a):
        n, wi, wo, wh = _v(n), _v(wi), _v(wo), _v(wh); return self.lib.l1_G_smith(_p(n), _p(wi), _p(wo), _p(wh), D(a))

    def microFacetProb(self, wo, wh, a, n):
        wo, wh, n = _v(wo), _v(wh), _v(n); return self.lib.l1_microFacetProb(_p(wo), _p(wh), D(a), _p(n))

    def frMicroFacet(self, eta, kappa, wi, wh, wo, a, n):
        e, k, wi, wh, wo, n = _v(eta), _v(kappa), _v(wi), _v(wh), _v(wo), _v(n); o = np.zeros(3)
        self.lib.l1_frMicroFacet(_p(e), _p(k), _p(wi), _p(wh), _p(wo), D(a), _p(n), _p(o)); return o

    def fresnelDie(self, ei, et, ct, ci): return self.lib.l1_fresnelDie(D(ei), D(et), D(ct), D(ci))
    def powerHeuristics(self, f, g): return self.lib.l1_powerHeuristics(D(f), D(g))

    def muestreoSA(self, scene, quirks, light, x, obj, n, wray, alpha, u):
        s, x, n, wray = _v(scene), _v(x), _v(n), _v(wray); u, pu, nu = self._u(u); Lo = np.zeros(3); wi = np.zeros(3); cm = D(0)
        self.lib.l1_muestreoSA(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(light), _p(x), C.c_int(obj), _p(n), _p(wray), D(alpha), pu, nu, _p(Lo), _p(wi), C.byref(cm))
        return Lo, wi, cm.value

    def MISv2(self, scene, quirks, obj, x, n, wray, alpha, st, u):
        s, x, n, wray = _v(scene), _v(x), _v(n), _v(wray); u, pu, nu = self._u(u); o = np.zeros(3)
        self.lib.l1_MISv2(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(obj), _p(x), _p(n), _p(wray), D(alpha), D(st), pu, nu, _p(o)); return o

    def pLight(self, scene, quirks, obj, x, n, wray, I, light, alpha):
        s, x, n, wray, I, light = _v(scene), _v(x), _v(n), _v(wray), _v(I), _v(light); o = np.zeros(3)
        self.lib.l1_pLight(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(obj), _p(x), _p(n), _p(wray), _p(I), _p(light), D(alpha), _p(o)); return o

    def bdsf(self, scene, wray, n, idx, u):
        s, wray, n = _v(scene), _v(wray), _v(n); u, pu, nu = self._u(u); wi = np.zeros(3); fs = np.zeros(3); pr = D(0)
        self.lib.l1_bdsf(_p(s), C.c_int(len(s)), _p(wray), _p(n), C.c_int(idx), pu, nu, _p(wi), C.byref(pr), _p(fs)); return fs, wi, pr.value

    # -- estimators
    def radiance_list(self, scene, quirks, method, sa, ss, o, d, u, cp=0.6, max_depth=0):
        s, o, d = _v(scene), _v(o), _v(d); u, pu, nu = self._u(u); out = np.zeros(3)
        used = self.lib.l1_radiance_list(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(method), D(sa), D(ss), D(cp), C.c_int(max_depth), _p(o), _p(d), pu, nu, _p(out))
        return out, used

    def radiance_erand48(self, scene, quirks, method, sa, ss, o, d, seed3, cp=0.6, max_depth=0):
        s, o, d = _v(scene), _v(o), _v(d); out = np.zeros(3)
        n = self.lib.l1_radiance_erand48(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(method), D(sa), D(ss), D(cp), C.c_int(max_depth), _p(o), _p(d),
                                         C.c_uint(seed3[0]), C.c_uint(seed3[1]), C.c_uint(seed3[2]), _p(out))
        return out, n

    def radiance_philox(self, scene, quirks, method, sa, ss, seed, o, d, pixel, sample, cp=0.6, max_depth=0):
        s = _v(scene); o = _v(o).reshape(-1, 3); d = _v(d).reshape(-1, 3); n = len(o)
        px = np.ascontiguousarray(pixel, dtype=np.uint32); sm = np.ascontiguousarray(sample, dtype=np.uint32)
        out = np.zeros((n, 3)); ev = np.zeros(n, dtype=np.uint64)
        self.lib.l1_radiance_philox(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(method), D(sa), D(ss), D(cp), C.c_int(max_depth), C.c_uint64(seed),
                                    C.c_int(n), _p(o), _p(d), px.ctypes.data_as(PU32), sm.ctypes.data_as(PU32), _p(out), ev.ctypes.data_as(PU64))
        return out, ev

    def camera(self, w, h, cam_o=CAM_O, cam_dir=CAM_DIR, fov=CAM_FOV):
        co, cd = _v(cam_o), _v(cam_dir); o, d, cx, cy = (np.zeros(3) for _ in range(4))
        self.lib.l1_camera(C.c_int(w), C.c_int(h), _p(co), _p(cd), D(fov), _p(o), _p(d), _p(cx), _p(cy)); return o, d, cx, cy

    def camera_ray(self, w, h, x, y, xi1, xi2, cam_o=CAM_O, cam_dir=CAM_DIR, fov=CAM_FOV):
        co, cd = _v(cam_o), _v(cam_dir); o = np.zeros(3)
        self.lib.l1_camera_ray(C.c_int(w), C.c_int(h), _p(co), _p(cd), D(fov), C.c_int(x), C.c_int(y), D(xi1), D(xi2), _p(o)); return o

    def render(self, scene, quirks, method, sa, ss, w, h, seed, spp, sample_begin=0, cp=0.6, max_depth=0, cam_o=CAM_O, cam_dir=CAM_DIR,
               fov=CAM_FOV, nthreads=0, want_sumsq=True):
        s, co, cd = _v(scene), _v(cam_o), _v(cam_dir)
        total = np.zeros((h, w, 3)); sq = np.zeros((h, w, 3)) if want_sumsq else None; st = np.zeros(3, dtype=np.uint64)
        self.lib.l1_render(_p(s), C.c_int(len(s)), C.c_uint(quirks), C.c_int(method), D(sa), D(ss), D(cp), C.c_int(max_depth), C.c_int(w), C.c_int(h),
                           _p(co), _p(cd), D(fov), C.c_uint64(seed), C.c_int(sample_begin), C.c_int(sample_begin + spp), C.c_int(nthreads),
                           _p(total), _p(sq) if want_sumsq else None, st.ctypes.data_as(PU64))
        return total, sq, dict(paths=int(st[0]), events=int(st[1]), scans=int(st[2]))

    def ray_march3(self, scene, quirks, sa, ss, step, source, o, d):
        """rayMarching3 on n rays -> n x 4 (L[3], steps)"""
        s = _v(scene); o = _v(o).reshape(-1, 3); d = _v(d).reshape(-1, 3); out = np.zeros((len(o), 4))
        self.lib.l1_ray_march3(_p(s), C.c_int(len(s)), C.c_uint(quirks), D(sa), D(ss), D(step), C.c_int(source), C.c_int(len(o)), _p(o), _p(d), _p(out))
        return out

    def render_march(self, scene, quirks, sa, ss, step, source, w, h, seed, spp, sample_begin=0, cam_o=CAM_O, cam_dir=CAM_DIR, fov=CAM_FOV, nthreads=0):
        s, co, cd = _v(scene), _v(cam_o), _v(cam_dir)
        total = np.zeros((h, w, 3))
        self.lib.l1_render_march(_p(s), C.c_int(len(s)), C.c_uint(quirks), D(sa), D(ss), D(step), C.c_int(source), C.c_int(w), C.c_int(h), _p(co), _p(cd), D(fov),
                                 C.c_uint64(seed), C.c_int(sample_begin), C.c_int(sample_begin + spp), C.c_int(nthreads), _p(total))
        return total

    def toDisplayValue(self, x): return self.lib.l1_toDisplayValue(D(x))


class L0:
    """The unmodified reference behind oracle/l0_harness.cpp."""

    def __init__(self, path=L0_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        for name in ("l0_erand48", "l0_sphere_intersect", "l0_cosinethetaMax", "l0_clamp", "l0_transmitance", "l0_freeFlightSample", "l0_freeFlightProb",
                     "l0_pdfSuccess", "l0_pdfFailure", "l0_isotropicPhaseProb", "l0_equiAngularSample", "l0_equiAngularProb", "l0_equiAngularParams2",
                     "l0_solidAngleProb", "l0_hemiCosineProb", "l0_NDF", "l0_G_smith", "l0_microFacetProb", "l0_fresnelDie", "l0_powerHeuristics"):
            getattr(L, name).restype = D
        L.l0_draws.restype = C.c_uint64
        L.l0_render.restype = C.c_uint64
        self._keep = None

    @staticmethod
    def available():
        return os.path.exists(L0_PATH)

    def seed(self, s): self.lib.l0_seed(C.c_uint(s[0]), C.c_uint(s[1]), C.c_uint(s[2]))
    def erand48(self): return self.lib.l0_erand48()
    def draws(self): return self.lib.l0_draws()

    def inject(self, u):
        self._keep = _v(u); self.lib.l0_inject(_p(self._keep), C.c_int(len(self._keep)))

    def set_hooks(self, robust_visibility, skip_r0): self.lib.l0_set_hooks(C.c_int(int(robust_visibility)), C.c_int(int(skip_r0)))

    def set_quirks(self, quirks):
        self.set_hooks(not (quirks & QUIRK_EXACT_VISIBILITY), not (quirks & QUIRK_R0_FALLTHROUGH))

    def scene(self):
        n = self.lib.l0_scene_size(); out = np.zeros((n, 18))
        for i in range(n):
            self.lib.l0_scene_get(C.c_int(i), _p(out[i]))
        return out

    def set_scene(self, scene):
        s = _v(scene); self.lib.l0_scene_set(C.c_int(len(s)), _p(s))

    def reset_scene(self): self.lib.l0_scene_reset()

    def sphere_intersect(self, i, o, d):
        o, d = _v(o), _v(d); return self.lib.l0_sphere_intersect(C.c_int(i), _p(o), _p(d))

    def intersect(self, o, d, id0=0):
        o, d = _v(o), _v(d); t = D(0); i = C.c_int(id0)
        h = self.lib.l0_intersect(_p(o), _p(d), C.byref(t), C.byref(i)); return bool(h), t.value, i.value

    def visibility(self, light, x):
        a, b = _v(light), _v(x); return bool(self.lib.l0_visibility(_p(a), _p(b)))

    def cosinethetaMax(self, i, x):
        x = _v(x); return self.lib.l0_cosinethetaMax(C.c_int(i), _p(x))

    def coordinateSystem(self, n):
        n = _v(n); s = np.zeros(3); t = np.zeros(3); self.lib.l0_coordinateSystem(_p(n), _p(s), _p(t)); return s, t

    def coordinateTraspose(self, n, w):
        n = _v(n); w = _v(w).copy(); self.lib.l0_coordinateTraspose(_p(n), _p(w)); return w

    def toDisplayValue(self, x): return self.lib.l0_toDisplayValue(D(x))
    def transmitance(self, a, b, st):
        a, b = _v(a), _v(b); return self.lib.l0_transmitance(_p(a), _p(b), D(st))

    def freeFlightSample(self, st, u): self.inject(u); return self.lib.l0_freeFlightSample(D(st))
    def freeFlightProb(self, st, d): return self.lib.l0_freeFlightProb(D(st), D(d))
    def pdfSuccess(self, st, t): return self.lib.l0_pdfSuccess(D(st), D(t))
    def pdfFailure(self, st, t): return self.lib.l0_pdfFailure(D(st), D(t))

    def isotropicPhaseSample(self, u):
        self.inject(u); o = np.zeros(3); self.lib.l0_isotropicPhaseSample(_p(o)); return o

    def equiAngularSample(self, Dd, a, b, u): self.inject(u); return self.lib.l0_equiAngularSample(D(Dd), D(a), D(b))
    def equiAngularProb(self, Dd, a, b, t): return self.lib.l0_equiAngularProb(D(Dd), D(a), D(b), D(t))

    def equiAngularParams2(self, src, tmax, o, d, u):
        self.inject(u); o, d = _v(o), _v(d); out = np.zeros(4)
        r = self.lib.l0_equiAngularParams2(C.c_int(src), D(tmax), _p(o), _p(d), _p(out)); return r, out

    def freeSingleScattering(self, xt, src, st, pS, u):
        self.inject(u); xt = _v(xt); o = np.zeros(3); self.lib.l0_freeSingleScattering(_p(xt), C.c_int(src), D(st), D(pS), _p(o)); return o

    def singleScattering(self, xt, src, st, ss, T, pS, u):
        self.inject(u); xt = _v(xt); o = np.zeros(3); self.lib.l0_singleScattering(_p(xt), C.c_int(src), D(st), D(ss), D(T), D(pS), _p(o)); return o

    def cosineHemispheric(self, n, u):
        self.inject(u); n = _v(n); o = np.zeros(3); self.lib.l0_cosineHemispheric(_p(n), _p(o)); return o

    def solidAngleDir(self, wc, cmax, u):
        self.inject(u); wc = _v(wc); o = np.zeros(3); self.lib.l0_solidAngleDir(_p(wc), D(cmax), _p(o)); return o

    def solidAngleProb(self, c): return self.lib.l0_solidAngleProb(D(c))
    def hemiCosineProb(self, c): return self.lib.l0_hemiCosineProb(D(c))

    def vectorFacet(self, alpha, u):
        self.inject(u); o = np.zeros(3); self.lib.l0_vectorFacet(D(alpha), _p(o)); return o

    def NDF(self, c, a): return self.lib.l0_NDF(D(c), D(a))

    def fresnel(self, c, eta, kappa):
        e, k = _v(eta), _v(kappa); o = np.zeros(3); self.lib.l0_fresnel(D(c), _p(e), _p(k), _p(o)); return o

    def G_smith(self, n, wi, wo, wh, a):
        n, wi, wo, wh = _v(n), _v(wi), _v(wo), _v(wh); return self.lib.l0_G_smith(_p(n), _p(wi), _p(wo), _p(wh), D(a))

    def microFacetProb(self, wo, wh, a, n):
        wo, wh, n = _v(wo), _v(wh), _v(n); return self.lib.l0_microFacetProb(_p(wo), _p(wh), D(a), _p(n))

    def frMicroFacet(self, eta, kappa, wi, wh, wo, a, n):
        e, k, wi, wh, wo, n = _v(eta), _v(kappa), _v(wi), _v(wh), _v(wo), _v(n); o = np.zeros(3)
        self.lib.l0_frMicroFacet(_p(e), _p(k), _p(wi), _p(wh), _p(wo), D(a), _p(n), _p(o)); return o

    def fresnelDie(self, ei, et, ct, ci): return self.lib.l0_fresnelDie(D(ei), D(et), D(ct), D(ci))

    def reflexDielectric(self, wi, n):
        wi, n = _v(wi), _v(n); o = np.zeros(3); self.lib.l0_reflexDielectric(_p(wi), _p(n), _p(o)); return o

    def refraxDielectric(self, ei, et, wi, n):
        wi, n = _v(wi), _v(n); o = np.zeros(3); self.lib.l0_refraxDielectric(D(ei), D(et), _p(wi), _p(n), _p(o)); return o

    def powerHeuristics(self, f, g): return self.lib.l0_powerHeuristics(D(f), D(g))

    def muestreoSA(self, light, x, obj, n, wray, alpha, u):
        self.inject(u); x, n, wray = _v(x), _v(n), _v(wray); Lo = np.zeros(3); wi = np.zeros(3); cm = D(0)
        self.lib.l0_muestreoSA(C.c_int(light), _p(x), C.c_int(obj), _p(n), _p(wray), D(alpha), _p(Lo), _p(wi), C.byref(cm)); return Lo, wi, cm.value

    def MISv2(self, obj, x, n, wray, alpha, st, u):
        self.inject(u); x, n, wray = _v(x), _v(n), _v(wray); o = np.zeros(3)
        self.lib.l0_MISv2(C.c_int(obj), _p(x), _p(n), _p(wray), D(alpha), D(st), _p(o)); return o

    def pLight(self, obj, x, n, wray, I, light, alpha):
        x, n, wray, I, light = _v(x), _v(n), _v(wray), _v(I), _v(light); o = np.zeros(3)
        self.lib.l0_pLight(C.c_int(obj), _p(x), _p(n), _p(wray), _p(I), _p(light), D(alpha), _p(o)); return o

    def bdsf(self, wray, n, idx, u):
        self.inject(u); wray, n = _v(wray), _v(n); wi = np.zeros(3); fs = np.zeros(3); pr = D(0)
        self.lib.l0_bdsf(_p(wray), _p(n), C.c_int(idx), _p(wi), C.byref(pr), _p(fs)); return fs, wi, pr.value

    def radiance(self, method, o, d, sa, ss, seed3=None, u=None):
        if u is not None:
            self.inject(u)
        elif seed3 is not None:
            self.inject([]); self.seed(seed3)
        o, d = _v(o), _v(d); out = np.zeros(3)
        self.lib.l0_radiance(C.c_int(method), _p(o), _p(d), D(sa), D(ss), _p(out)); return out, self.draws()

    def rayMarching3(self, o, d, sa, ss, step, src):
        o, d = _v(o), _v(d); out = np.zeros(3)
        self.lib.l0_rayMarching3(_p(o), _p(d), D(sa), D(ss), D(step), C.c_int(src), _p(out)); return out

    def camera(self, w, h):
        o, d, cx, cy = (np.zeros(3) for _ in range(4)); self.lib.l0_camera(C.c_int(w), C.c_int(h), _p(o), _p(d), _p(cx), _p(cy)); return o, d, cx, cy

    def camera_ray(self, w, h, x, y, xi1, xi2):
        o = np.zeros(3); self.lib.l0_camera_ray(C.c_int(w), C.c_int(h), C.c_int(x), C.c_int(y), D(xi1), D(xi2), _p(o)); return o

    def render(self, w, h, spp, method, sa, ss, seed, nthreads=0, want_sumsq=True):
        total = np.zeros((h, w, 3)); sq = np.zeros((h, w, 3)) if want_sumsq else None
        draws = self.lib.l0_render(C.c_int(w), C.c_int(h), C.c_int(spp), C.c_int(method), D(sa), D(ss), C.c_uint64(seed), C.c_int(nthreads),
                                   _p(total), _p(sq) if want_sumsq else None)
        return total, sq, int(draws)
