"""ctypes binding of the C oracle (oracle/sia2d_c.c) -- TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libsia2d_oracle.so")


class _Par64(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("dx", "dy", "eta0", "n", "p", "q", "rho", "g", "C", "A")] + [("Afield", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            import sys

            sys.path.insert(0, os.path.dirname(_HERE))
            import __graft_entry__ as ge

            ge.build_oracle()
        _lib = C.CDLL(SO)
        _lib.sia2d_oracle_threads.restype = C.c_int
    return _lib


def threads():
    return int(lib().sia2d_oracle_threads())


def use_all_cores():
    """Run the C oracle on every host core available to this process, whatever OMP_NUM_THREADS says (torchrun sets it
    to 1).  Returns the thread count now in effect."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().sia2d_oracle_set_threads(C.c_int(n))
    return threads()


def _par(dx, dy, ph, A, npdt):
    p = _Par64(dx, dy, ph.eta0, ph.n, ph.p, ph.q, ph.rho, ph.g, ph.C, 0.0, None)
    keep = None
    if np.ndim(A) == 2:
        keep = np.asfortranarray(A, dtype=npdt)
        p.Afield = keep.ctypes.data
    else:
        p.A = float(A)
    return p, keep


def rhs(H, B, dx, dy, ph, A, dtype=np.float64):
    """F1 on one glacier; arrays are (nx, ny), returned column-major."""
    H = np.asfortranarray(H, dtype=dtype)
    B = np.asfortranarray(B, dtype=dtype)
    nx, ny = H.shape
    out = np.empty((nx, ny), dtype=dtype, order="F")
    work = np.empty((nx - 1) * (ny - 1), dtype=dtype)
    p, keep = _par(dx, dy, ph, A, dtype)
    fn = lib().sia2d_rhs_f64 if dtype == np.float64 else lib().sia2d_rhs_f32
    fn(C.c_int(nx), C.c_int(ny), C.c_void_p(H.ctypes.data), C.c_void_p(B.ctypes.data), C.c_void_p(out.ctypes.data),
       C.c_void_p(work.ctypes.data), C.byref(p))
    return out


def vjp(lam, H, B, dx, dy, ph, A, dtype=np.float64, want_H=True, want_field=False):
    """A1 + A2 on one glacier: returns (vjp_H or None, S, vjpA field or None)."""
    H = np.asfortranarray(H, dtype=dtype)
    B = np.asfortranarray(B, dtype=dtype)
    lam = np.asfortranarray(lam, dtype=dtype)
    nx, ny = H.shape
    out = np.empty((nx, ny), dtype=dtype, order="F") if want_H else None
    fld = np.empty((nx - 1, ny - 1), dtype=dtype, order="F") if want_field else None
    work = np.empty(4 * (nx - 1) * (ny - 1), dtype=dtype)
    S = C.c_double(0.0)
    p, keep = _par(dx, dy, ph, A, dtype)
    fn = lib().sia2d_vjp_f64 if dtype == np.float64 else lib().sia2d_vjp_f32
    fn(C.c_int(nx), C.c_int(ny), C.c_void_p(lam.ctypes.data), C.c_void_p(H.ctypes.data), C.c_void_p(B.ctypes.data),
       C.c_void_p(out.ctypes.data if want_H else None), C.byref(S), C.c_void_p(fld.ctypes.data if want_field else None),
       C.c_void_p(work.ctypes.data), C.byref(p))
    return out, S.value, fld


def solve_fixed(H0, B, dx, dy, ph, A, t, nsub=8, method="ssprk3", dtype=np.float64):
    """oracle.sia2d_numpy.solve_forward("euler" | "ssprk3") as one C loop: returns the (n_t, nx, ny) snapshots
    (each [j] a column-major (nx, ny) matrix)."""
    H0 = np.asfortranarray(H0, dtype=dtype)
    B = np.asfortranarray(B, dtype=dtype)
    nx, ny = H0.shape
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.empty((t.size, ny, nx), dtype=dtype)  # plane j, C order (ny, nx) == column-major (nx, ny)
    work = np.empty((nx - 1) * (ny - 1) + 3 * nx * ny, dtype=dtype)
    p, keep = _par(dx, dy, ph, A, dtype)
    fn = lib().sia2d_solve_fixed_f64 if dtype == np.float64 else lib().sia2d_solve_fixed_f32
    fn(C.c_int(nx), C.c_int(ny), C.c_void_p(H0.ctypes.data), C.c_void_p(B.ctypes.data), C.byref(p), C.c_int(t.size),
       C.c_void_p(t.ctypes.data), C.c_int(nsub), C.c_int({"euler": 0, "ssprk3": 1}[method]), C.c_void_p(out.ctypes.data),
       C.c_void_p(work.ctypes.data))
    return [out[j].T for j in range(t.size)]


def grad_discrete(B, dx, dy, ph, A, t, Hs, Href, masks, dtype=np.float64):
    """oracle.sia2d_numpy.loss_and_grad_discrete for a glacier-wide law as one C loop: returns (loss, Ssum, lambda(t_0))
    with d(theta) = (dA/dtheta) * Ssum."""
    B = np.asfortranarray(B, dtype=dtype)
    nx, ny = B.shape
    t = np.ascontiguousarray(t, dtype=np.float64)
    pack = lambda arrs: np.ascontiguousarray(np.stack([np.asarray(a, dtype=dtype).T for a in arrs]))
    hs, hr = pack(Hs), pack(Href)
    w = pack([np.asarray(m, dtype=np.float64) / float(nx * ny) for m in masks])
    lam = np.empty((ny, nx), dtype=dtype)
    work = np.empty(4 * (nx - 1) * (ny - 1) + nx * ny, dtype=dtype)
    loss, Ssum = C.c_double(0.0), C.c_double(0.0)
    p, keep = _par(dx, dy, ph, A, dtype)
    fn = lib().sia2d_grad_discrete_f64 if dtype == np.float64 else lib().sia2d_grad_discrete_f32
    fn(C.c_int(nx), C.c_int(ny), C.c_void_p(B.ctypes.data), C.byref(p), C.c_int(t.size), C.c_void_p(t.ctypes.data),
       C.c_void_p(hs.ctypes.data), C.c_void_p(hr.ctypes.data), C.c_void_p(w.ctypes.data), C.byref(loss), C.byref(Ssum),
       C.c_void_p(lam.ctypes.data), C.c_void_p(work.ctypes.data))
    return loss.value, Ssum.value, lam.T
