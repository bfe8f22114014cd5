"""cfd-codes_b200 -- B200-native (sm_100a) lid-driven-cavity Lattice-Boltzmann hot path.

Python-side mirror of the C ABI in ``include/lbm_b200.h`` (ctypes; plain pointers and sizes).  The
product is the CUDA library ``liblbm_b200.so`` built from ``csrc/``; this module only loads it and
gives the entry points the names the reference program uses
(`/root/reference/3D Lid-driven Cavity with Lattice Boltzmann Method/main.c`, "3D:" below):

    reference (main.c)                               here
    -----------------------------------------------  -----------------------------------------
    allocate(); initialize();          3D:135, 172   Cavity3D(N=..., Re=..., ...)
    macrocollideandstream(); HechtBC(f2); swap(...)  Cavity3D.step(n)              3D:106-123
    diff(f2, f1, NQ)                       3D:678    Cavity3D.residual()
    f2 (global array, LU4 order)           3D:48,70  Cavity3D.f()
    macrovars(f2); calcvmag()          3D:695, 730   Cavity3D.macrovars()
    datwrite(...)                          3D:363    Cavity3D.datwrite(path)
    freemem()                              3D:150    Cavity3D.close()

There is NO CPU fallback: if the CUDA library is missing or no GPU is present the constructor
raises.  Because the directory name contains a hyphen, import it through ``lbmpkg.py`` at the repo
root (``from lbmpkg import lbm``).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblbm_b200.so")

LBM_OK, LBM_ERR_INVALID, LBM_ERR_CUDA, LBM_ERR_UNSTABLE, LBM_ERR_NCCL, LBM_ERR_NOMEM = range(6)
NCCL_ID_BYTES = 128
PEER_BLOB_BYTES = 320

# every symbol include/lbm_b200.h declares
ABI_SYMBOLS = [
    "lbm_default_params", "lbm_nccl_unique_id", "lbm_create", "lbm_peer_export", "lbm_peer_connect", "lbm_step", "lbm_step_timed", "lbm_timer_start",
    "lbm_timer_stop", "lbm_residual", "lbm_converge",
    "lbm_get_f", "lbm_set_f", "lbm_set_f_async", "lbm_set_f_commit", "lbm_macrovars", "lbm_macrovars_async", "lbm_io_wait", "lbm_sync", "lbm_get_info", "lbm_set_option", "lbm_destroy",
    "lbm_last_error", "lbm_version",
]


class LbmParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int), ("N", C.c_int), ("M", C.c_int), ("Re", C.c_double), ("UMAX", C.c_double),
        ("METHOD", C.c_int), ("GAMMA", C.c_double), ("MAGIC", C.c_double), ("DIAG", C.c_int),
        ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int), ("track_residual", C.c_int),
        ("stream", C.c_void_p), ("nccl_id", C.c_ubyte * NCCL_ID_BYTES),
    ]


class LbmInfo(C.Structure):
    _fields_ = [
        ("NU", C.c_double), ("TAU", C.c_double), ("OMEGA", C.c_double), ("OMEGAm", C.c_double),
        ("OmOMEGA", C.c_double), ("ulid", C.c_double), ("vlid", C.c_double), ("wlid", C.c_double),
        ("Q", C.c_int), ("k0", C.c_int), ("k1", C.c_int), ("cells_local", C.c_longlong),
        ("cells_global", C.c_longlong), ("steps_done", C.c_longlong), ("kernel_launches", C.c_longlong),
        ("device_bytes", C.c_longlong), ("plain_step_ms", C.c_double), ("plain_step_iterations", C.c_longlong),
        ("pair_enabled", C.c_int), ("pair_variant", C.c_int),
    ]


class LbmBlockRecord(C.Structure):
    _fields_ = [("iteration", C.c_longlong), ("df", C.c_double), ("status", C.c_int), ("last", C.c_int)]


BLOCK_FN = C.CFUNCTYPE(None, C.POINTER(LbmBlockRecord), C.c_void_p)


class LbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lbm_b200 error %d: %s" % (code, msg))
        self.code = code


class UnstableError(LbmError):
    """The reference's `Error: negative feq.  Therefore, unstable.` + exit(1) (3D:687-692)."""


_lib = None


def load_library(path=None):
    """Load liblbm_b200.so (the CUDA product).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not built: run `make -C cfd-codes_b200` (or __graft_entry__.build()); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, dp = C.c_void_p, C.c_void_p
    L.lbm_default_params.argtypes = [C.POINTER(LbmParams)]
    L.lbm_default_params.restype = None
    L.lbm_nccl_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.lbm_create.argtypes = [C.POINTER(LbmParams), C.POINTER(vp)]
    L.lbm_peer_export.argtypes = [vp, C.POINTER(C.c_ubyte)]
    L.lbm_peer_connect.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.lbm_step.argtypes = [vp, C.c_longlong]
    L.lbm_step_timed.argtypes = [vp, C.c_longlong, C.POINTER(C.c_float)]
    L.lbm_timer_start.argtypes = [vp]
    L.lbm_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.lbm_residual.argtypes = [vp, C.POINTER(C.c_double)]
    L.lbm_converge.argtypes = [vp, C.c_longlong, C.c_double, C.c_longlong, C.c_double, BLOCK_FN, C.c_void_p,
                               C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    L.lbm_get_f.argtypes = [vp, dp]
    L.lbm_set_f.argtypes = [vp, dp, dp]
    L.lbm_macrovars.argtypes = [vp, dp, dp, dp, dp, dp]
    L.lbm_set_f_async.argtypes = [vp, dp, dp]
    L.lbm_set_f_commit.argtypes = [vp]
    L.lbm_macrovars_async.argtypes = [vp, dp, dp, dp, dp, dp]
    L.lbm_io_wait.argtypes = [vp]
    L.lbm_sync.argtypes = [vp]
    L.lbm_get_info.argtypes = [vp, C.POINTER(LbmInfo)]
    L.lbm_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.lbm_destroy.argtypes = [vp]
    L.lbm_last_error.restype = C.c_char_p
    L.lbm_version.restype = C.c_char_p
    for name in ("lbm_nccl_unique_id", "lbm_create", "lbm_peer_export", "lbm_peer_connect", "lbm_step", "lbm_step_timed", "lbm_timer_start", "lbm_timer_stop",
                 "lbm_residual", "lbm_converge", "lbm_get_f", "lbm_set_f_async", "lbm_set_f_commit", "lbm_macrovars_async", "lbm_io_wait",
                 "lbm_set_f", "lbm_macrovars", "lbm_sync", "lbm_get_info", "lbm_set_option", "lbm_destroy"):
        getattr(L, name).restype = C.c_int
    if path == LIB_PATH:
        _lib = L
    return L


def _check(rc):
    if rc != LBM_OK:
        msg = load_library().lbm_last_error().decode()
        raise (UnstableError if rc == LBM_ERR_UNSTABLE else LbmError)(rc, msg)


def default_params():
    p = LbmParams()
    load_library().lbm_default_params(C.byref(p))
    return p


def nccl_unique_id():
    buf = (C.c_ubyte * NCCL_ID_BYTES)()
    _check(load_library().lbm_nccl_unique_id(buf))
    return bytes(buf)


def slab_range(n, rank, nranks):
    """Planes of the slowest axis owned by `rank` (same formula as csrc/lbm_b200.cu:create3d)."""
    return rank * n // nranks, (rank + 1) * n // nranks


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else int(a))


class Cavity3D:
    """One context = one GPU = one z-slab of the N^3 D3Q19 cavity (parameters as 3D:7-15, 22)."""

    Q = 19

    def __init__(self, N=50, Re=100, UMAX=0.1, METHOD=1, GAMMA=1.0, MAGIC=0.25, DIAG=0, rank=0, nranks=1,
                 device=-1, track_residual=True, nccl_id=None, stream=None):
        self._h = None
        L = load_library()
        p = default_params()
        p.dim, p.N, p.Re, p.UMAX, p.METHOD, p.GAMMA, p.MAGIC, p.DIAG = 3, N, Re, UMAX, METHOD, GAMMA, MAGIC, DIAG
        p.rank, p.nranks, p.device, p.track_residual = rank, nranks, device, 1 if track_residual else 0
        p.stream = stream
        if nranks > 1:
            if nccl_id is None or len(nccl_id) != NCCL_ID_BYTES:
                raise ValueError("nranks > 1 needs the shared nccl_unique_id() bytes")
            C.memmove(p.nccl_id, nccl_id, NCCL_ID_BYTES)
        self.params = p
        self.N = N
        h = C.c_void_p()
        _check(L.lbm_create(C.byref(p), C.byref(h)))
        self._h = h
        self._L = L

    # -- direct NVLink halo (optional) ----------------------------------------------------------------
    def peer_export(self):
        buf = (C.c_ubyte * PEER_BLOB_BYTES)()
        _check(self._L.lbm_peer_export(self._h, buf))
        return bytes(buf)

    def peer_connect(self, below, above):
        _check(self._L.lbm_peer_connect(self._h, below, above))

    # -- the reference's time loop ---------------------------------------------------------------
    def step(self, n=1):
        _check(self._L.lbm_step(self._h, n))

    def step_timed(self, n):
        """n iterations, device time in ms (CUDA events on the launching stream)."""
        ms = C.c_float()
        _check(self._L.lbm_step_timed(self._h, n, C.byref(ms)))
        return ms.value

    def timer_start(self):
        _check(self._L.lbm_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        _check(self._L.lbm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def residual(self):
        out = C.c_double()
        _check(self._L.lbm_residual(self._h, C.byref(out)))
        return out.value

    def converge(self, skip, thresh, max_blocks, df0, on_block=None):
        """The reference's convergence loop (3D:103-124) as ONE call: up to max_blocks blocks of `skip`
        iterations, df = diff * df0 after each, stop when df < thresh.  On one GPU the loop runs on
        the device (conditional CUDA graph) without host synchronisation between the blocks.
        Returns (blocks_done, last df, [(iteration, df), ...])."""
        recs = []

        def cb(rec, _user):
            r = rec.contents
            recs.append((r.iteration, r.df))
            if on_block:
                on_block(r.iteration, r.df)

        fn = BLOCK_FN(cb)
        done, df = C.c_longlong(), C.c_double()
        _check(self._L.lbm_converge(self._h, skip, thresh, max_blocks, df0, fn, None, C.byref(done), C.byref(df)))
        return done.value, df.value, recs

    def sync(self):
        _check(self._L.lbm_sync(self._h))

    # -- state access -------------------------------------------------------------------------------
    def f(self, out=None):
        """The reference's f2 (LU4 order, NQ doubles); only this rank's planes are written."""
        if out is None:
            out = np.zeros(self.Q * self.N ** 3, dtype=np.float64)
        _check(self._L.lbm_get_f(self._h, _ptr(out)))
        return out

    def set_f(self, f, f_prev=None):
        f = np.ascontiguousarray(f, dtype=np.float64)
        if f_prev is not None:
            f_prev = np.ascontiguousarray(f_prev, dtype=np.float64)
        _check(self._L.lbm_set_f(self._h, _ptr(f), _ptr(f_prev)))

    def macrovars(self, out=None):
        n3 = self.N ** 3
        if out is None:
            out = {k: np.zeros(n3, dtype=np.float64) for k in ("u1", "u2", "u3", "rho", "vmag")}
        _check(self._L.lbm_macrovars(self._h, _ptr(out["u1"]), _ptr(out["u2"]), _ptr(out["u3"]), _ptr(out["rho"]),
                                     _ptr(out["vmag"])))
        return out

    def raw_set_f(self, f_ptr, f_prev_ptr=None):
        _check(self._L.lbm_set_f(self._h, C.c_void_p(f_ptr), C.c_void_p(f_prev_ptr) if f_prev_ptr else None))

    # pipelined host I/O (pinned host buffers given as raw addresses)
    def raw_set_f_async(self, f_ptr, f_prev_ptr=None):
        _check(self._L.lbm_set_f_async(self._h, C.c_void_p(f_ptr), C.c_void_p(f_prev_ptr) if f_prev_ptr else None))

    def set_f_commit(self):
        _check(self._L.lbm_set_f_commit(self._h))

    def raw_macrovars_async(self, ptrs):
        _check(self._L.lbm_macrovars_async(self._h, *[C.c_void_p(p) if p else None for p in ptrs]))

    def io_wait(self):
        _check(self._L.lbm_io_wait(self._h))

    def raw_macrovars(self, ptrs):
        _check(self._L.lbm_macrovars(self._h, *[C.c_void_p(p) if p else None for p in ptrs]))

    def info(self):
        i = LbmInfo()
        _check(self._L.lbm_get_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in LbmInfo._fields_}

    def set_option(self, key, value):
        _check(self._L.lbm_set_option(self._h, key.encode(), int(value)))

    # -- output (3D:363-382) ------------------------------------------------------------------------
    def filename(self):
        p = self.params
        return "Solution_n=%dRe=%d_DIAG=%d_%dRT_M=%.3f.dat" % (p.N, int(p.Re), p.DIAG, p.METHOD + 1, p.MAGIC)  # 3D:174

    def datwrite(self, path=None, mv=None):
        """Tecplot POINT file in the reference's format: i outer, j, k inner, `%.10f` x 7 (3D:372-376)."""
        mv = mv or self.macrovars()
        N = self.N
        name = self.filename()
        path = path or name
        u = mv["u1"].reshape(N, N, N)  # [k][j][i]
        v = mv["u2"].reshape(N, N, N)
        w = mv["u3"].reshape(N, N, N)
        m = mv["vmag"].reshape(N, N, N)
        with open(path, "w") as fh:
            fh.write('TITLE="%s" VARIABLES="x", "y", "z", "u", "v", "w", "vmag" ZONE T="%s" I=%d J=%d K=%d F=POINT\n'
                     % (name, name, N, N, N))
            kk, jj = np.meshgrid(np.arange(N, dtype=np.float64), np.arange(N, dtype=np.float64))   # rows: j outer, k inner
            for i in range(N):     # one i-slab (N^2 lines) per formatting call
                slab = np.column_stack([np.full(N * N, float(i)), jj.ravel(), kk.ravel(), u[:, :, i].T.ravel(),
                                        v[:, :, i].T.ravel(), w[:, :, i].T.ravel(), m[:, :, i].T.ravel()])
                np.savetxt(fh, slab, fmt="%.10f", delimiter=", ")
        return path

    def close(self):
        if self._h:
            self._L.lbm_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_to_convergence(sim, THRESH=1e-12, skip=500, MAXITER=int(1e8), log=None):
    """The reference's driver loop (3D:93-124): iteration 0 normalises (df0 = 1/diff), then every
    `skip` iterations df = diff*df0 and the loop stops when df < THRESH.  Returns (iterations, df).
    The blocks run inside Cavity3D.converge (device-side loop on one GPU)."""
    sim.step(1)
    df0 = sim.residual()
    if log:
        log("\nIteration 0:\ndf:\t%.3e" % df0)
    df0 = 1.0 / df0
    blocks = max(0, (MAXITER - 1) // skip)
    done, df, _ = sim.converge(skip, THRESH, blocks, df0,
                               (lambda t, d: log("\nIteration %d:\ndf/df0:\t%.3e" % (t, d))) if log else None)
    return done * skip, (df if done else None)


class Cavity2D:
    """D2Q9 N x M cavity, the variant of the same kernel named by BASELINE config 5; parameters
    as the reference's 2D program (2D main.cpp:12-17, 25): Re, N (width), M (height), Ulid, MAGIC.

        reference (main.cpp)                                   here
        -----------------------------------------------------  ----------------------------
        one sweep of streamCollide over all rows   2D:160-178  Cavity2D.step(n)
        macroVars(...) + convergence(...)          2D:419-478  Cavity2D.residual()
        dist (LU3 layout 9N*j + N*k + i)           2D:276-278  Cavity2D.dist()
        u, v, rho arrays / outputScalar            2D:482-495  Cavity2D.macro(), Cavity2D.output_bin()

    The update is the time-consistent two-lattice version of the reference's in-place sweep
    (DESIGN.md 7); residual() is the change of (u, v, rho) since the previous residual() call."""

    Q = 9

    def __init__(self, N=100, M=300, Re=100.0, Ulid=0.1, MAGIC=0.25, rank=0, nranks=1, device=-1, nccl_id=None):
        self._h = None
        L = load_library()
        p = default_params()
        p.dim, p.N, p.M, p.Re, p.UMAX, p.METHOD, p.MAGIC = 2, N, M, Re, Ulid, 1, MAGIC
        p.rank, p.nranks, p.device, p.track_residual = rank, nranks, device, 0
        if nranks > 1:
            if nccl_id is None or len(nccl_id) != NCCL_ID_BYTES:
                raise ValueError("nranks > 1 needs the shared nccl_unique_id() bytes")
            C.memmove(p.nccl_id, nccl_id, NCCL_ID_BYTES)
        self.params, self.N, self.M = p, N, M
        h = C.c_void_p()
        _check(L.lbm_create(C.byref(p), C.byref(h)))
        self._h, self._L = h, L

    def step(self, n=1):
        _check(self._L.lbm_step(self._h, n))

    def step_timed(self, n):
        ms = C.c_float()
        _check(self._L.lbm_step_timed(self._h, n, C.byref(ms)))
        return ms.value

    def residual(self):
        out = C.c_double()
        _check(self._L.lbm_residual(self._h, C.byref(out)))
        return out.value

    def sync(self):
        _check(self._L.lbm_sync(self._h))

    def dist(self, out=None):
        if out is None:
            out = np.zeros(9 * self.N * self.M, dtype=np.float64)
        _check(self._L.lbm_get_f(self._h, _ptr(out)))
        return out

    def set_dist(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        _check(self._L.lbm_set_f(self._h, _ptr(d), None))

    def macro(self, out=None):
        n = self.N * self.M
        if out is None:
            out = {k: np.zeros(n, dtype=np.float64) for k in ("u", "v", "rho", "vmag")}
        _check(self._L.lbm_macrovars(self._h, _ptr(out["u"]), _ptr(out["v"]), None, _ptr(out["rho"]), _ptr(out.get("vmag"))))
        return out

    def info(self):
        i = LbmInfo()
        _check(self._L.lbm_get_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in LbmInfo._fields_}

    def set_option(self, key, value):
        _check(self._L.lbm_set_option(self._h, key.encode(), int(value)))

    def output_bin(self, directory=".", mv=None):
        """The reference's output files `Re=%.0f_N=%d_M=%d_{x,y,u,v,rho}.bin` (2D:190-202, 482-495):
        raw little-endian doubles, M rows x N columns; x = 0.5 + i, y = 0.5 + j (2D:146-152)."""
        mv = mv or self.macro()
        p = self.params
        j, i = np.meshgrid(np.arange(self.M, dtype=np.float64), np.arange(self.N, dtype=np.float64), indexing="ij")
        fields = dict(x=0.5 + i, y=0.5 + j, u=mv["u"], v=mv["v"], rho=mv["rho"])
        paths = []
        for name, a in fields.items():
            path = os.path.join(directory, "Re=%.0f_N=%d_M=%d_%s.bin" % (p.Re, p.N, p.M, name))
            np.ascontiguousarray(a, dtype="<f8").tofile(path)
            paths.append(path)
        return paths

    def close(self):
        if self._h:
            self._L.lbm_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_to_convergence_2d(sim, THRESH=2e-9, prntInt=1000, MAXITER=int(1e6), log=None):
    """The 2D reference's loop (2D:160-187, 452-478): macroVars/convergence after iterations
    1, prntInt+1, 2*prntInt+1, ...; df0 from the first check.  Returns (t of the last check, df)."""
    sim.step(1)
    d0 = sim.residual()
    if log:
        log("Iteration\tdf/df0\t\tMLUPS")
        log("%.3e\t%.3e" % (0.0, 1.0))
    t, df = 0, 1.0
    while t + prntInt < MAXITER:
        sim.step(prntInt)
        t += prntInt
        df = sim.residual() / d0
        if log:
            log("%.3e\t%.3e" % (float(t), df))
        if df < THRESH:
            break
    return t, df
