"""ctypes binding of the CPU oracle (oracle/libdem_oracle.so) and, when built, of the reference's own
objects (oracle/_ref/libchrono_ref.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (chrono_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdem_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libchrono_ref.so")

HOOKE, HERTZ, PLAINCOULOMB, FLORES = 0, 1, 2, 3
ADH_CONSTANT, ADH_DMT, ADH_PERKO = 0, 1, 2
TANG_NONE, TANG_ONESTEP, TANG_MULTISTEP = 0, 1, 2
SHAPE_SPHERE, SHAPE_BOX, SHAPE_TRIANGLE = 0, 2, 13


class OrcMaterial(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "young", "poisson", "mu_s", "mu_roll", "mu_spin", "cr", "adhesion", "adhesion_dmt", "adhesion_perko",
        "kn", "kt", "gn", "gt")]


class OrcSettings(C.Structure):
    _fields_ = [("force_model", C.c_int), ("adhesion_model", C.c_int), ("tangential_mode", C.c_int),
                ("use_mat_props", C.c_int), ("char_vel", C.c_double), ("min_slip_vel", C.c_double),
                ("min_roll_vel", C.c_double), ("min_spin_vel", C.c_double), ("dt", C.c_double),
                ("gravity", C.c_double * 3), ("bins_per_axis", C.c_int * 3), ("num_threads", C.c_int)]


def make_material(young=2e5, poisson=0.3, mu_s=0.6, mu_roll=0.0, mu_spin=0.0, cr=0.4, adhesion=0.0,
                  adhesion_dmt=0.0, adhesion_perko=0.0, kn=2e5, kt=2e5, gn=40.0, gt=20.0):
    """Defaults follow ChContactMaterialSMC (src/chrono/physics/ChContactMaterialSMC.cpp:28-38) and
    ChContactMaterial (mu 0.6, cr 0.4)."""
    return OrcMaterial(young, poisson, mu_s, mu_roll, mu_spin, cr, adhesion, adhesion_dmt, adhesion_perko,
                       kn, kt, gn, gt)


def make_settings(force_model=HERTZ, adhesion_model=ADH_CONSTANT, tangential_mode=TANG_MULTISTEP,
                  use_mat_props=True, char_vel=1.0, min_slip_vel=1e-4, min_roll_vel=1e-4, min_spin_vel=1e-4,
                  dt=1e-3, gravity=(0, 0, -9.81), bins=(10, 10, 10), num_threads=0):
    """Defaults: src/chrono_multicore/ChSettings.h:117-124 (tangential mode set explicitly)."""
    s = OrcSettings()
    s.force_model, s.adhesion_model, s.tangential_mode = force_model, adhesion_model, tangential_mode
    s.use_mat_props = int(use_mat_props)
    s.char_vel, s.min_slip_vel, s.min_roll_vel, s.min_spin_vel = char_vel, min_slip_vel, min_roll_vel, min_spin_vel
    s.dt = dt
    s.gravity[:] = gravity
    s.bins_per_axis[:] = bins
    s.num_threads = num_threads
    return s


def build(ref=True, quiet=True):
    """Compile the oracle (always) and the reference objects (only where /root/reference exists)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "-j8"], stdout=out)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "-j8", "ref"], stdout=out)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build(ref=False)
        L = C.CDLL(LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcSettings)]
        for name in ("orc_destroy", "orc_set_settings", "orc_set_body_state", "orc_set_body_fixed",
                     "orc_get_body_state", "orc_get_grid", "orc_get_shape_bins", "orc_get_broadphase_sizes",
                     "orc_get_bin_csr", "orc_get_pairs", "orc_get_contacts", "orc_get_contact_forces",
                     "orc_get_body_forces", "orc_get_history", "orc_add_history", "orc_get_timers",
                     "orc_reset_timers", "orc_composite", "orc_generate_aabb"):
            getattr(L, name).restype = None
        L.orc_num_contacts.restype = C.c_longlong
        L.orc_num_history.restype = C.c_longlong
        L.orc_snap_to_box.restype = C.c_uint
        _lib = L
    return _lib


class Oracle:
    """One Multicore-SMC system.  Body/shape numbering follows the reference: bodies in
    insertion order, shapes in global insertion order."""

    def __init__(self, settings):
        self.L = lib()
        self.settings = settings
        self.h = C.c_void_p(self.L.orc_create(C.byref(settings)))

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_settings(self, s):
        self.settings = s
        self.L.orc_set_settings(self.h, C.byref(s))

    def add_material(self, m):
        return self.L.orc_add_material(self.h, C.byref(m))

    def add_body(self, mass, inertia, pos, rot=(1, 0, 0, 0), vel=(0, 0, 0), omega=(0, 0, 0), fixed=False):
        a = [_f64(x) for x in (inertia, pos, rot, vel, omega)]
        return self.L.orc_add_body(self.h, C.c_double(mass), _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]),
                                   int(fixed))

    def add_spheres(self, pos, radius, mass, material, vel=None, omega=None):
        pos = _f64(pos, (-1, 3))
        n = pos.shape[0]
        radius = _f64(np.broadcast_to(radius, (n,)))
        mass = _f64(np.broadcast_to(mass, (n,)))
        vel = _f64(vel, (-1, 3)) if vel is not None else np.zeros((n, 3))
        omega = _f64(omega, (-1, 3)) if omega is not None else np.zeros((n, 3))
        return self.L.orc_add_spheres(self.h, n, _dp(pos), _dp(vel), _dp(omega), _dp(radius), _dp(mass), material)

    def add_box(self, body, material, lpos, hdims, lrot=(1, 0, 0, 0)):
        a = [_f64(x) for x in (lpos, lrot, hdims)]
        return self.L.orc_add_box(self.h, body, material, _dp(a[0]), _dp(a[1]), _dp(a[2]))

    def add_triangles(self, body, material, verts):
        v = _f64(verts, (-1, 9))
        return self.L.orc_add_triangles(self.h, body, material, v.shape[0], _dp(v))

    def set_body_state(self, body, pos=None, rot=None, vel=None, omega=None):
        a = [(_f64(x) if x is not None else None) for x in (pos, rot, vel, omega)]
        self.L.orc_set_body_state(self.h, body, *[(_dp(x) if x is not None else None) for x in a])

    def step(self, n=1):
        return self.L.orc_step(self.h, n)

    def eval(self):
        return self.L.orc_eval(self.h)

    @property
    def num_bodies(self):
        return self.L.orc_num_bodies(self.h)

    @property
    def num_shapes(self):
        return self.L.orc_num_shapes(self.h)

    def state(self):
        nb = self.num_bodies
        pos, rot, vel, om = np.empty((nb, 3)), np.empty((nb, 4)), np.empty((nb, 3)), np.empty((nb, 3))
        self.L.orc_get_body_state(self.h, _dp(pos), _dp(rot), _dp(vel), _dp(om))
        return pos, rot, vel, om

    def generate_aabb(self):
        ns = self.num_shapes
        mn, mx = np.empty((ns, 3)), np.empty((ns, 3))
        self.L.orc_generate_aabb(self.h, _dp(mn), _dp(mx))
        return mn, mx

    def grid(self):
        o, b, ib = np.empty(3), np.empty(3), np.empty(3)
        bins = np.empty(3, dtype=np.int32)
        self.L.orc_get_grid(self.h, _dp(o), _dp(b), _dp(ib), _ip(bins))
        return o, b, ib, bins

    def shape_bins(self):
        ns = self.num_shapes
        gmin, gmax = np.empty((ns, 3), dtype=np.int32), np.empty((ns, 3), dtype=np.int32)
        self.L.orc_get_shape_bins(self.h, _ip(gmin), _ip(gmax))
        return gmin, gmax

    def broadphase_sizes(self):
        s = (C.c_longlong * 3)()
        self.L.orc_get_broadphase_sizes(self.h, s)
        return tuple(s)

    def bin_csr(self):
        nab, nint, _ = self.broadphase_sizes()
        act, start, num = (np.empty(nab, dtype=np.uint32), np.empty(nab + 1, dtype=np.uint32),
                           np.empty(nint, dtype=np.uint32))
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint))
        self.L.orc_get_bin_csr(self.h, up(act), up(start), up(num))
        return act, start, num

    def pairs(self):
        n = self.broadphase_sizes()[2]
        p = np.empty(n, dtype=np.int64)
        self.L.orc_get_pairs(self.h, p.ctypes.data_as(C.POINTER(C.c_longlong)))
        return p

    def contacts(self):
        n = self.L.orc_num_contacts(self.h)
        out = dict(shape_pair=np.empty(n, dtype=np.int64), body_pair=np.empty((n, 2), dtype=np.int32),
                   normal=np.empty((n, 3)), depth=np.empty(n), pt1=np.empty((n, 3)), pt2=np.empty((n, 3)),
                   erad=np.empty(n))
        self.L.orc_get_contacts(self.h, out["shape_pair"].ctypes.data_as(C.POINTER(C.c_longlong)),
                                _ip(out["body_pair"]), _dp(out["normal"]), _dp(out["depth"]), _dp(out["pt1"]),
                                _dp(out["pt2"]), _dp(out["erad"]))
        return out

    def contact_forces(self):
        n = self.L.orc_num_contacts(self.h)
        f, t1, t2 = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        self.L.orc_get_contact_forces(self.h, _dp(f), _dp(t1), _dp(t2))
        return f, t1, t2

    def body_forces(self):
        nb = self.num_bodies
        f, t = np.zeros((nb, 3)), np.zeros((nb, 3))
        self.L.orc_get_body_forces(self.h, _dp(f), _dp(t))
        return f, t

    def history(self):
        n = self.L.orc_num_history(self.h)
        out = dict(body=np.empty(n, dtype=np.int32), other=np.empty(n, dtype=np.int32),
                   shape1=np.empty(n, dtype=np.int32), shape2=np.empty(n, dtype=np.int32), disp=np.empty((n, 3)),
                   duration=np.empty(n), relvel_init=np.empty(n))
        self.L.orc_get_history(self.h, _ip(out["body"]), _ip(out["other"]), _ip(out["shape1"]), _ip(out["shape2"]),
                               _dp(out["disp"]), _dp(out["duration"]), _dp(out["relvel_init"]))
        return out

    def add_history(self, body, other, shape1, shape2, disp, duration=0.0, relvel_init=0.0):
        d = _f64(disp)
        self.L.orc_add_history(self.h, body, other, shape1, shape2, _dp(d), C.c_double(duration),
                               C.c_double(relvel_init))

    def timers(self):
        t = (C.c_double * 5)()
        self.L.orc_get_timers(self.h, t)
        return dict(broad=t[0], narrow=t[1], force=t[2], integrate=t[3], total=t[4])

    def reset_timers(self):
        self.L.orc_reset_timers(self.h)


# ---------------------------------------------------------------------------------------------------------
# primitive-level entry points (same call shape for the restatement "orc" and the reference "ref")
# ---------------------------------------------------------------------------------------------------------
def _prim_call(fn, *args):
    norm, pt1, pt2 = np.zeros(3), np.zeros(3), np.zeros(3)
    depth, erad = C.c_double(0), C.c_double(0)
    hit = fn(*args, _dp(norm), C.byref(depth), _dp(pt1), _dp(pt2), C.byref(erad))
    if not hit:
        return None
    return dict(norm=norm, depth=depth.value, pt1=pt1, pt2=pt2, erad=erad.value)


class Prims:
    """sphere_sphere / box_sphere / triangle_sphere / snap_* through either library (prefix 'orc' or 'ref')."""

    def __init__(self, L, prefix):
        self.L, self.p = L, prefix
        getattr(L, prefix + "_snap_to_box").restype = C.c_uint

    def sphere_sphere(self, pos1, r1, pos2, r2, sep=0.0):
        a, b = _f64(pos1), _f64(pos2)
        return _prim_call(getattr(self.L, self.p + "_sphere_sphere"), _dp(a), C.c_double(r1), _dp(b), C.c_double(r2),
                          C.c_double(sep))

    def box_sphere(self, pos1, rot1, hdims1, pos2, r2, sep=0.0):
        a, q, h, b = _f64(pos1), _f64(rot1), _f64(hdims1), _f64(pos2)
        return _prim_call(getattr(self.L, self.p + "_box_sphere"), _dp(a), _dp(q), _dp(h), _dp(b), C.c_double(r2),
                          C.c_double(sep))

    def triangle_sphere(self, A, B, Cc, pos2, r2, sep=0.0):
        a, b, c, p = _f64(A), _f64(B), _f64(Cc), _f64(pos2)
        return _prim_call(getattr(self.L, self.p + "_triangle_sphere"), _dp(a), _dp(b), _dp(c), _dp(p),
                          C.c_double(r2), C.c_double(sep))

    def snap_to_box(self, hdims, loc):
        h, l = _f64(hdims), _f64(loc).copy()
        code = getattr(self.L, self.p + "_snap_to_box")(_dp(h), _dp(l))
        return code, l

    def snap_to_triangle(self, A, B, Cc, P):
        a, b, c, p = _f64(A), _f64(B), _f64(Cc), _f64(P)
        r = np.zeros(3)
        edge = getattr(self.L, self.p + "_snap_to_triangle")(_dp(a), _dp(b), _dp(c), _dp(p), _dp(r))
        return bool(edge), r


def orc_prims():
    return Prims(lib(), "orc")


def rotate(v, q):
    v, q = _f64(v), _f64(q)
    a, b, c = np.zeros(3), np.zeros(3), np.zeros(3)
    lib().orc_rotate.restype = None
    lib().orc_rotate(_dp(v), _dp(q), _dp(a), _dp(b), _dp(c))
    return a, b, c


def composite(m1, m2):
    out = np.zeros(13)
    lib().orc_composite(C.byref(m1), C.byref(m2), _dp(out))
    return out


def contact_force(which, settings, comp, b1, b2, mass, pos, rot, vel, pt1, pt2, normal, depth, erad, hist=None):
    """One contact through the force law; which = 'orc' (restatement) or 'ref' (reference object code).
    hist = None | dict(disp, dur, relvel).  Returns (force_on_b2, torque_b1, torque_b2, hist_out)."""
    mass, pos, rot, vel = _f64(mass), _f64(pos), _f64(rot), _f64(vel)
    pt1, pt2, normal, comp = _f64(pt1), _f64(pt2), _f64(normal), _f64(comp)
    present = C.c_int(0 if hist is None else 1)
    hd = _f64(hist["disp"]).copy() if hist else np.zeros(3)
    hdur = C.c_double(hist["dur"] if hist else 0.0)
    hrel = C.c_double(hist["relvel"] if hist else 0.0)
    F, T1, T2 = np.zeros(3), np.zeros(3), np.zeros(3)
    if which == "orc":
        lib().orc_contact_force(C.byref(settings), _dp(comp), b1, b2, _dp(mass), _dp(pos), _dp(rot), _dp(vel),
                                _dp(pt1), _dp(pt2), _dp(normal), C.c_double(depth), C.c_double(erad),
                                C.byref(present), _dp(hd), C.byref(hdur), C.byref(hrel), _dp(F), _dp(T1), _dp(T2))
    else:
        model4 = np.array([settings.force_model, settings.adhesion_model, settings.tangential_mode,
                           settings.use_mat_props], dtype=np.int32)
        par5 = np.array([settings.char_vel, settings.min_slip_vel, settings.min_roll_vel, settings.min_spin_vel,
                         settings.dt])
        ref().L.ref_contact_force(_ip(model4), _dp(par5), _dp(comp), b1, b2, _dp(mass), _dp(pos), _dp(rot), _dp(vel),
                                  _dp(pt1), _dp(pt2), _dp(normal), C.c_double(depth), C.c_double(erad),
                                  C.byref(present), _dp(hd), C.byref(hdur), C.byref(hrel), _dp(F), _dp(T1), _dp(T2))
    hout = dict(disp=hd, dur=hdur.value, relvel=hrel.value) if present.value else None
    return F, T1, T2, hout


class Ref:
    """The reference's own broadphase/narrowphase/force objects (oracle/_ref/libchrono_ref.so)."""

    def __init__(self):
        self.L = C.CDLL(REF_PATH)
        for name in ("ref_rotate", "ref_collision_run", "ref_cd_sizes", "ref_cd_grid", "ref_cd_bins", "ref_cd_pairs",
                     "ref_cd_contacts"):
            getattr(self.L, name).restype = None
        self.prims = Prims(self.L, "ref")

    def rotate(self, v, q):
        v, q = _f64(v), _f64(q)
        a, b, c = np.zeros(3), np.zeros(3), np.zeros(3)
        self.L.ref_rotate(_dp(v), _dp(q), _dp(a), _dp(b), _dp(c))
        return a, b, c

    def collision(self, types, bodies, lpos, lrot, dims, tri, pos, rot, active, collide, aabb_min, aabb_max, bins):
        types = np.ascontiguousarray(types, dtype=np.int32)
        bodies = np.ascontiguousarray(bodies, dtype=np.int32)
        ns, nb = len(types), len(pos)
        lpos, lrot, dims, tri = _f64(lpos, (ns, 3)), _f64(lrot, (ns, 4)), _f64(dims, (ns, 3)), _f64(tri, (ns, 9))
        pos, rot = _f64(pos, (nb, 3)), _f64(rot, (nb, 4))
        active = np.ascontiguousarray(active, dtype=np.int8)
        collide = np.ascontiguousarray(collide, dtype=np.int8)
        aabb_min, aabb_max = _f64(aabb_min, (ns, 3)), _f64(aabb_max, (ns, 3))
        bins = np.ascontiguousarray(bins, dtype=np.int32)
        cp = lambda a: a.ctypes.data_as(C.c_char_p)
        self.L.ref_collision_run(ns, _ip(types), _ip(bodies), _dp(lpos), _dp(lrot), _dp(dims), _dp(tri), nb, _dp(pos),
                                 _dp(rot), cp(active), cp(collide), _dp(aabb_min), _dp(aabb_max), _ip(bins))
        s = (C.c_longlong * 4)()
        self.L.ref_cd_sizes(s)
        nab, nint, npair, nc = tuple(s)
        o, b, ib = np.empty(3), np.empty(3), np.empty(3)
        self.L.ref_cd_grid(_dp(o), _dp(b), _dp(ib))
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint))
        act, start, num = (np.empty(nab, dtype=np.uint32), np.empty(nab + 1, dtype=np.uint32),
                           np.empty(nint, dtype=np.uint32))
        self.L.ref_cd_bins(up(act), up(start), up(num))
        pairs = np.empty(npair, dtype=np.int64)
        self.L.ref_cd_pairs(pairs.ctypes.data_as(C.POINTER(C.c_longlong)))
        ct = dict(shape_pair=np.empty(nc, dtype=np.int64), body_pair=np.empty((nc, 2), dtype=np.int32),
                  normal=np.empty((nc, 3)), depth=np.empty(nc), pt1=np.empty((nc, 3)), pt2=np.empty((nc, 3)),
                  erad=np.empty(nc))
        self.L.ref_cd_contacts(ct["shape_pair"].ctypes.data_as(C.POINTER(C.c_longlong)), _ip(ct["body_pair"]),
                               _dp(ct["normal"]), _dp(ct["depth"]), _dp(ct["pt1"]), _dp(ct["pt2"]), _dp(ct["erad"]))
        return dict(origin=o, bin_size=b, inv_bin_size=ib, bin_active=act, bin_start_index=start,
                    bin_aabb_number=num, pairs=pairs, contacts=ct)


_ref = None


def ref_available():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref
