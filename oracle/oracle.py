"""oracle/oracle.py -- TEST INFRASTRUCTURE: ctypes access to the CPU oracle and to oracle/_ref.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package (cfd-codes_b200) never does.
"""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblbm_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

Q3 = 19
# lattice velocities of the reference, main.c:36-38
C3 = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
               [1, 1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, -1], [0, 1, 1], [0, -1, -1],
               [1, -1, 0], [-1, 1, 0], [1, 0, -1], [-1, 0, 1], [0, 1, -1], [0, -1, 1]], dtype=np.int64)
W3 = np.array([1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12)


class Params3D(C.Structure):
    _fields_ = [("N", C.c_int), ("Re", C.c_int), ("METHOD", C.c_int), ("DIAG", C.c_int),
                ("UMAX", C.c_double), ("GAMMA", C.c_double), ("MAGIC", C.c_double)]


_lib = None


def build():
    """(Re)build liblbm_oracle.so with oracle/Makefile."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liblbm_oracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle3d_create.restype = C.c_void_p
        L.oracle3d_create.argtypes = [C.POINTER(Params3D)]
        L.oracle3d_destroy.argtypes = [C.c_void_p]
        L.oracle3d_step.restype = C.c_int
        L.oracle3d_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle3d_last_diff.restype = C.c_double
        L.oracle3d_last_diff.argtypes = [C.c_void_p]
        L.oracle3d_unstable.restype = C.c_int
        L.oracle3d_unstable.argtypes = [C.c_void_p]
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.oracle3d_get_f.argtypes = [C.c_void_p, dp]
        L.oracle3d_get_f_prev.argtypes = [C.c_void_p, dp]
        L.oracle3d_derived.argtypes = [C.c_void_p, dp]
        L.oracle3d_macrovars.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.oracle3d_time_steps.restype = C.c_double
        L.oracle3d_time_steps.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


class Oracle3D:
    """CPU restatement of the reference 3D solver's time loop (oracle/lbm3d_oracle.c)."""

    def __init__(self, N=50, Re=100, UMAX=0.1, METHOD=1, GAMMA=1.0, MAGIC=0.25, DIAG=0, threads=0):
        self.N, self.threads = N, threads
        self.p = Params3D(N=N, Re=Re, METHOD=METHOD, DIAG=DIAG, UMAX=UMAX, GAMMA=GAMMA, MAGIC=MAGIC)
        self.h = lib().oracle3d_create(C.byref(self.p))
        if not self.h:
            raise MemoryError("oracle3d_create failed")

    def step(self, n=1, want_diff=True):
        return lib().oracle3d_step(self.h, n, self.threads, 1 if want_diff else 0)

    @property
    def last_diff(self):
        return lib().oracle3d_last_diff(self.h)

    @property
    def unstable(self):
        return bool(lib().oracle3d_unstable(self.h))

    def f(self):
        out = np.empty(Q3 * self.N ** 3, dtype=np.float64)
        lib().oracle3d_get_f(self.h, out)
        return out

    def f_prev(self):
        out = np.empty(Q3 * self.N ** 3, dtype=np.float64)
        lib().oracle3d_get_f_prev(self.h, out)
        return out

    def derived(self):
        out = np.empty(8, dtype=np.float64)
        lib().oracle3d_derived(self.h, out)
        return dict(zip(["NU", "TAU", "OMEGA", "OMEGAm", "OmOMEGA", "ulid", "vlid", "wlid"], out.tolist()))

    def macrovars(self):
        n3 = self.N ** 3
        arrs = [np.empty(n3, dtype=np.float64) for _ in range(5)]
        lib().oracle3d_macrovars(self.h, *arrs)
        return dict(zip(["u1", "u2", "u3", "rho", "vmag"], arrs))

    def time_steps(self, n):
        return lib().oracle3d_time_steps(self.h, n, self.threads)

    def close(self):
        if self.h:
            lib().oracle3d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the real reference, compiled under oracle/_ref by build_ref.py -------------------------

def ref_tag(N=50, Re=100, METHOD=1, GAMMA="1.0", DIAG=0, UMAX="0.1"):
    return "N{}_Re{}_m{}_g{}_d{}_u{}".format(N, Re, METHOD, GAMMA, DIAG, UMAX)


def ref_binary(**cfg):
    return os.path.join(REF_DIR, "lbm3d_ref_" + ref_tag(**cfg))


def ref_available(**cfg):
    return os.path.exists(ref_binary(**cfg))


def ref_steps(n, threads=0, **cfg):
    """Run n iterations of the real reference; returns (f [NQ doubles, LU4 order], last_diff,
    returncode).  returncode 1 == the reference's own `negative feq` exit (main.c:687-692)."""
    exe = ref_binary(**cfg)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "f.bin")
        r = subprocess.run([exe, "steps", str(n), out, str(threads)], capture_output=True, text=True, cwd=tmp)
        if r.returncode != 0:
            return None, None, r.returncode, r.stdout
        info = json.loads(r.stdout.strip().splitlines()[-1])
        return np.fromfile(out, dtype=np.float64), info["last_diff"], 0, r.stdout


def ref_time(n, threads=0, **cfg):
    exe = ref_binary(**cfg)
    r = subprocess.run([exe, "time", str(n), str(threads)], capture_output=True, text=True, check=True)
    return json.loads(r.stdout.strip().splitlines()[-1])


# ---- D2Q9 (2D main.cpp) ---------------------------------------------------------------------------

class Params2D(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("Re", C.c_double), ("Ulid", C.c_double), ("MAGIC", C.c_double)]


def _lib2d():
    L = lib()
    if not getattr(L, "_has2d", False):
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        for pre in ("oracle2d", "inplace2d"):
            getattr(L, pre + "_create").restype = C.c_void_p
            getattr(L, pre + "_destroy").argtypes = [C.c_void_p]
            getattr(L, pre + "_step").argtypes = [C.c_void_p, C.c_int]
            getattr(L, pre + "_residual").restype = C.c_double
            getattr(L, pre + "_residual").argtypes = [C.c_void_p]
            getattr(L, pre + "_get_macro").argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle2d_create.argtypes = [C.POINTER(Params2D)]
        L.inplace2d_create.argtypes = [C.POINTER(Params2D), C.c_int]
        L.oracle2d_get_dist.argtypes = [C.c_void_p, dp]
        L.oracle2d_derived.argtypes = [C.POINTER(Params2D), dp]
        L._has2d = True
    return L


class Oracle2D:
    """D2Q9 cavity.  scheme='sync': two-lattice synchronous update (what the GPU computes);
    scheme='inplace': literal emulation of the reference's in-place sweep with `ranks` MPI ranks."""

    def __init__(self, N=100, M=300, Re=100.0, Ulid=0.1, MAGIC=0.25, scheme="sync", ranks=3):
        self.N, self.M, self.scheme = N, M, scheme
        self.p = Params2D(N=N, M=M, Re=Re, Ulid=Ulid, MAGIC=MAGIC)
        L = _lib2d()
        self.pre = "oracle2d" if scheme == "sync" else "inplace2d"
        self.h = L.oracle2d_create(C.byref(self.p)) if scheme == "sync" else L.inplace2d_create(C.byref(self.p), ranks)

    def step(self, n=1):
        getattr(_lib2d(), self.pre + "_step")(self.h, n)

    def residual(self):
        return getattr(_lib2d(), self.pre + "_residual")(self.h)

    def macro(self):
        a = [np.empty(self.N * self.M) for _ in range(3)]
        getattr(_lib2d(), self.pre + "_get_macro")(self.h, *a)
        return dict(u=a[0], v=a[1], rho=a[2])

    def dist(self):
        assert self.scheme == "sync"
        out = np.empty(9 * self.N * self.M)
        _lib2d().oracle2d_get_dist(self.h, out)
        return out

    def derived(self):
        out = np.empty(4)
        _lib2d().oracle2d_derived(C.byref(self.p), out)
        return dict(zip(["NU", "TAU", "OMEGA", "OMEGAm"], out.tolist()))

    def close(self):
        if self.h:
            getattr(_lib2d(), self.pre + "_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref2d_binary(N, M, Re, MAXITER, prntInt, THRESH):
    return os.path.join(REF_DIR, "lbm2d_ref_N{}_M{}_Re{}_it{}_p{}_th{}".format(N, M, Re, MAXITER, prntInt, THRESH))


def ref2d_run(ranks, **cfg):
    """Run the REAL 2D reference (built over oracle/mpi_shim) with `ranks` ranks in a temp dir;
    returns (dict of u, v, rho, x, y arrays [M*N], stdout)."""
    exe = ref2d_binary(**cfg)
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, SHIM_NP=str(ranks))
        r = subprocess.run([exe], capture_output=True, text=True, cwd=tmp, env=env, check=True)
        out = {}
        for name in ("u", "v", "rho", "x", "y"):
            out[name] = np.fromfile(os.path.join(tmp, "Re=%.0f_N=%d_M=%d_%s.bin" % (float(cfg["Re"]), cfg["N"], cfg["M"], name)))
        return out, r.stdout
