"""ctypes binding of the C ABI in include/chrono_b200_dem.h (libchrono_b200_dem.so).

This is plumbing for tests and bench.py: every call goes straight to the extern "C" entry points that a
Chrono maintainer would bind from C++.  Loading fails loudly when the library is missing --
there is no CPU fallback."""
import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

HOOKE, HERTZ, PLAINCOULOMB, FLORES = 0, 1, 2, 3
ADH_CONSTANT, ADH_DMT, ADH_PERKO = 0, 1, 2
TANG_NONE, TANG_ONESTEP, TANG_MULTISTEP = 0, 1, 2
FORWARD_EULER, CHUNG, CENTERED_DIFFERENCE, EXTENDED_TAYLOR = 0, 1, 2, 3
MAT_SPHERE, MAT_WALL, MAT_MESH = 0, 1, 2
RED_MAX_Z, RED_MIN_Z, RED_KE, RED_MAX_SPEED, RED_COUNT_ABOVE_Z, RED_COUNT_ABOVE_X, RED_NUM_CONTACTS, RED_KE_TRANSLATIONAL = range(8)


class Material(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "young", "poisson", "mu_s", "mu_roll", "mu_spin", "cr", "adhesion", "adhesion_dmt", "adhesion_perko",
        "kn", "kt", "gn", "gt")]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("force_model", C.c_int), ("adhesion_model", C.c_int),
                ("tangential_mode", C.c_int), ("use_mat_props", C.c_int), ("integrator", C.c_int),
                ("history_slots", C.c_int), ("char_vel", C.c_double), ("min_slip_vel", C.c_double),
                ("min_roll_vel", C.c_double), ("min_spin_vel", C.c_double), ("dt", C.c_double),
                ("gravity", C.c_double * 3), ("bins_per_axis", C.c_int * 3), ("material", Material * 3),
                ("mass_coef", C.c_double), ("wall_mass", C.c_double), ("mesh_mass", C.c_double),
                ("verlet_skin", C.c_double), ("neighbor_slots", C.c_int)]


def material(young=2e5, poisson=0.3, mu_s=0.6, mu_roll=0.0, mu_spin=0.0, cr=0.4, adhesion=0.0, adhesion_dmt=0.0,
             adhesion_perko=0.0, kn=2e5, kt=2e5, gn=40.0, gt=20.0):
    return Material(young, poisson, mu_s, mu_roll, mu_spin, cr, adhesion, adhesion_dmt, adhesion_perko, kn, kt, gn, gt)


def config(device=0, force_model=HERTZ, adhesion_model=ADH_CONSTANT, tangential_mode=TANG_MULTISTEP,
           use_mat_props=True, integrator=CENTERED_DIFFERENCE, history_slots=16, char_vel=1.0, min_slip_vel=1e-4,
           min_roll_vel=1e-4, min_spin_vel=1e-4, dt=1e-3, gravity=(0, 0, -9.81), bins=(10, 10, 10),
           mat_sphere=None, mat_wall=None, mat_mesh=None, mass_coef=4.0 / 3.0 * np.pi * 2000.0, wall_mass=1.0,
           mesh_mass=1.0, verlet_skin=-1.0, neighbor_slots=0):
    c = Config()
    c.device = device
    c.force_model, c.adhesion_model, c.tangential_mode = force_model, adhesion_model, tangential_mode
    c.use_mat_props, c.integrator, c.history_slots = int(use_mat_props), integrator, history_slots
    c.char_vel, c.min_slip_vel, c.min_roll_vel, c.min_spin_vel = char_vel, min_slip_vel, min_roll_vel, min_spin_vel
    c.dt = dt
    c.gravity[:] = gravity
    c.bins_per_axis[:] = bins
    ms = mat_sphere or material()
    c.material[MAT_SPHERE] = ms
    c.material[MAT_WALL] = mat_wall or ms
    c.material[MAT_MESH] = mat_mesh or ms
    c.mass_coef, c.wall_mass, c.mesh_mass = mass_coef, wall_mass, mesh_mass
    c.verlet_skin, c.neighbor_slots = verlet_skin, neighbor_slots
    return c


class DemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dem_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libchrono_b200_dem.so is not built (run __graft_entry__.build()); no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.dem_b200_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.dem_b200_destroy.restype = None
        L.dem_b200_destroy.argtypes = [C.c_void_p]
        L.dem_b200_last_error.restype = C.c_char_p
        L.dem_b200_last_error.argtypes = [C.c_void_p]
        L.dem_b200_kernel_name.restype = C.c_char_p
        L.dem_b200_num_spheres.restype = C.c_size_t
        L.dem_b200_num_spheres.argtypes = [C.c_void_p]
        L.dem_b200_time.restype = C.c_double
        L.dem_b200_time.argtypes = [C.c_void_p]
        L.dem_b200_num_walls.argtypes = [C.c_void_p]
        L.dem_b200_num_triangles.restype = C.c_size_t
        L.dem_b200_num_triangles.argtypes = [C.c_void_p]
        L.dem_b200_add_mesh.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_double]
        L.dem_b200_set_mesh_motion.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 4
        L.dem_b200_mesh_wrench.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


class DemSystem:
    def __init__(self, cfg):
        self.L = lib()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.dem_b200_create(C.byref(cfg), C.byref(h))
        if rc:
            raise DemError(rc, self.L.dem_b200_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.dem_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise DemError(rc, self.L.dem_b200_last_error(self.h).decode())
        return rc

    def set_config(self, cfg):
        self.cfg = cfg
        self._ck(self.L.dem_b200_set_config(self.h, C.byref(cfg)))

    def set_spheres(self, pos, radius, vel=None, omega=None, fixed=None):
        pos = _f64(pos, (-1, 3))
        n = pos.shape[0]
        radius = _f64(np.broadcast_to(radius, (n,)))
        vel, omega = _f64(vel, (n, 3)), _f64(omega, (n, 3))
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        self._ck(self.L.dem_b200_set_spheres(self.h, C.c_size_t(n), _dp(pos), _dp(vel), _dp(omega), _dp(radius),
                                             fx.ctypes.data_as(C.POINTER(C.c_uint8)) if fx is not None else None))
        self.n = n

    def add_box_wall(self, pos, hdims, rot=(1, 0, 0, 0)):
        p, q, h = _f64(pos), _f64(rot), _f64(hdims)
        return self._ck(self.L.dem_b200_add_box_wall(self.h, _dp(p), _dp(q), _dp(h)))

    def add_sphere_wall(self, center, radius, spheres_outside=True):
        c = _f64(center)
        return self._ck(self.L.dem_b200_add_sphere_wall(self.h, _dp(c), C.c_double(radius), int(spheres_outside)))

    def add_zcone_wall(self, tip, slope, hmin, hmax, spheres_above=True):
        t = _f64(tip)
        return self._ck(self.L.dem_b200_add_zcone_wall(self.h, _dp(t), C.c_double(slope), C.c_double(hmin), C.c_double(hmax),
                                                       int(spheres_above)))

    def add_plane_wall(self, pos, normal):
        p, nrm = _f64(pos), _f64(normal)
        return self._ck(self.L.dem_b200_add_plane_wall(self.h, _dp(p), _dp(nrm)))

    # ---- triangle meshes (ChSystemDemMesh) ----
    def add_mesh(self, verts9, mass=0.0):
        v = _f64(verts9, (-1, 9))
        return self._ck(self.L.dem_b200_add_mesh(self.h, C.c_size_t(v.shape[0]), _dp(v), C.c_double(mass)))

    def set_mesh_motion(self, mesh, pos=None, rot=None, lin_vel=None, ang_vel=None):
        a = [_f64(x) if x is not None else None for x in (pos, rot, lin_vel, ang_vel)]
        self._ck(self.L.dem_b200_set_mesh_motion(self.h, int(mesh), *[_dp(x) for x in a]))

    def enable_mesh_collision(self, enabled=True):
        self._ck(self.L.dem_b200_enable_mesh_collision(self.h, int(enabled)))

    def mesh_wrench(self, mesh):
        f, t = np.empty(3), np.empty(3)
        self._ck(self.L.dem_b200_mesh_wrench(self.h, int(mesh), _dp(f), _dp(t)))
        return f, t

    @property
    def num_triangles(self):
        return self.L.dem_b200_num_triangles(self.h)

    def initialize(self):
        self._ck(self.L.dem_b200_initialize(self.h))

    def step(self, n=1, sync=True):
        self._ck(self.L.dem_b200_step(self.h, int(n)))
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.L.dem_b200_sync(self.h))

    def step_timed(self, n):
        ms = C.c_float(0)
        self._ck(self.L.dem_b200_step_timed(self.h, int(n), C.byref(ms)))
        return ms.value

    def step_profile(self, n):
        ms = (C.c_float * 16)()
        k = C.c_int(0)
        self._ck(self.L.dem_b200_step_profile(self.h, int(n), ms, C.byref(k)))
        return {self.L.dem_b200_kernel_name(i).decode(): ms[i] for i in range(k.value)}

    def advance_host(self, pos, vel, omega, nsteps, pos_out, vel_out, omega_out):
        self._ck(self.L.dem_b200_advance_host(self.h, C.c_size_t(self.n), _dp(pos), _dp(vel), _dp(omega), int(nsteps),
                                              _dp(pos_out), _dp(vel_out), _dp(omega_out)))

    def state(self):
        n = self.n
        pos, vel, om = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        self._ck(self.L.dem_b200_get_state(self.h, _dp(pos), _dp(vel), _dp(om)))
        return pos, vel, om

    def set_state(self, pos=None, vel=None, omega=None):
        n = self.n
        pos, vel, omega = _f64(pos, (n, 3)), _f64(vel, (n, 3)), _f64(omega, (n, 3))
        self._ck(self.L.dem_b200_set_state(self.h, _dp(pos), _dp(vel), _dp(omega)))

    def request_rebuild(self):
        """The next step rebuilds the neighbour lists (no result depends on it)."""
        self._ck(self.L.dem_b200_request_rebuild(self.h))

    def sphere(self, i):
        p, v, w = np.empty(3), np.empty(3), np.empty(3)
        self._ck(self.L.dem_b200_get_sphere(self.h, C.c_size_t(i), _dp(p), _dp(v), _dp(w)))
        return p, v, w

    @property
    def time(self):
        return self.L.dem_b200_time(self.h)

    @property
    def num_walls(self):
        return self.L.dem_b200_num_walls(self.h)

    def reduce(self, which, arg=0.0):
        out = C.c_double(0)
        self._ck(self.L.dem_b200_reduce(self.h, which, C.c_double(arg), C.byref(out)))
        return out.value

    def enable_recording(self, enable=True, max_pairs=0):
        self._ck(self.L.dem_b200_enable_recording(self.h, int(enable), C.c_size_t(max_pairs)))

    def forces(self):
        f, t = np.empty((self.n, 3)), np.empty((self.n, 3))
        self._ck(self.L.dem_b200_get_forces(self.h, _dp(f), _dp(t)))
        return f, t

    def accel(self):
        """Linear acceleration of the last step, gravity included, user order (GetParticleLinAcc)."""
        a = np.empty((self.n, 3))
        self._ck(self.L.dem_b200_get_accel(self.h, _dp(a)))
        return a

    def pairs(self):
        n = C.c_size_t(0)
        self._ck(self.L.dem_b200_get_pairs(self.h, None, C.c_size_t(0), C.byref(n)))
        p = np.empty(n.value, dtype=np.uint64)
        if n.value:
            self._ck(self.L.dem_b200_get_pairs(self.h, p.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(n.value),
                                               C.byref(n)))
        return p

    def bins(self):
        gmin, gmax = np.empty((self.n, 3), dtype=np.int32), np.empty((self.n, 3), dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._ck(self.L.dem_b200_get_bins(self.h, ip(gmin), ip(gmax)))
        return gmin, gmax

    def grid(self):
        o, b, ib = np.empty(3), np.empty(3), np.empty(3)
        self._ck(self.L.dem_b200_get_grid(self.h, _dp(o), _dp(b), _dp(ib)))
        return o, b, ib

    def history(self):
        n = C.c_size_t(0)
        self._ck(self.L.dem_b200_get_history(self.h, None, None, None, None, None, C.c_size_t(0), C.byref(n)))
        m = n.value
        out = dict(owner=np.empty(m, dtype=np.uint32), other=np.empty(m, dtype=np.uint32), disp=np.empty((m, 3)),
                   duration=np.empty(m), relvel_init=np.empty(m))
        up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
        if m:
            self._ck(self.L.dem_b200_get_history(self.h, up(out["owner"]), up(out["other"]), _dp(out["disp"]),
                                                 _dp(out["duration"]), _dp(out["relvel_init"]), C.c_size_t(m),
                                                 C.byref(n)))
        return out

    def stats(self):
        a, b, c = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0)
        self._ck(self.L.dem_b200_get_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(steps=a.value, rebuilds=b.value, contacts_last_step=c.value)

    def add_history(self, owner_shape, other_shape, disp, duration=0.0, relvel_init=0.0):
        d = _f64(disp)
        self._ck(self.L.dem_b200_add_history(self.h, C.c_uint32(owner_shape), C.c_uint32(other_shape), _dp(d),
                                             C.c_double(duration), C.c_double(relvel_init)))
