"""gp.GP over several GPUs of one box: host-side mirror of the reference's Observe / Gradient / Absorb / LML
(gp/gp.go:374-413, 418-499, 80-87, 244-253) over the gogp_grid_* entry points.  The covariance matrix is dealt
2D block-cyclically to a pr x pc grid of GPUs and NCCL runs inside the library (gogp_b200/csrc/grid.hpp, grid.cu).

Two ways to form the grid, as in the C-ABI:

    g = GridGP(NDim=4, Simil=..., Noise=..., Devices=[0, 1, 2, 3])          # one process drives the GPUs
    g = GridGP(NDim=4, Simil=..., Noise=..., Rank=r, World=w, Device=d, UniqueId=id)   # SPMD, e.g. torchrun

``unique_id()`` makes the 128-byte id one rank creates and the host distributes (bench.py broadcasts it with
torch.distributed -- plumbing).  Results are replicated on every rank.
"""
import ctypes as C

import numpy as np

from . import _lib
from .gp import GoGPError, GoGPPanic, _flat


def unique_id():
    buf = (C.c_ubyte * _lib.GRID_ID_BYTES)()
    st = _lib.lib().gogp_grid_unique_id(buf)
    if st != _lib.OK:
        raise GoGPPanic(st, "gogp_grid_unique_id: NCCL is not available")
    return bytes(buf)


class GridGP:
    def __init__(self, NDim=1, Simil=None, Noise=None, Devices=None, Grid=(0, 0), Block=0, Rank=None, World=None,
                 Device=0, UniqueId=None):
        self.NDim, self.Simil, self.Noise = NDim, Simil, Noise
        self.ThetaSimil, self.ThetaNoise = [], []
        self.X, self.Y = [], []
        self._g = C.c_void_p()
        self._data_key = None
        L = _lib.lib()
        if Simil is None:
            raise GoGPPanic(_lib.BAD_ARGUMENT, "GP.Simil is nil")
        sd = Simil.Descriptor()
        if Noise is not None:
            nd, nn, ntn = Noise.Descriptor(), len(Noise.Descriptor()), Noise.NTheta()
        else:
            nd, nn, ntn = None, 0, 0
        pr, pc = Grid
        if Rank is None:
            devs = list(Devices if Devices is not None else [0])
            arr = (C.c_int * len(devs))(*devs)
            st = L.gogp_create_grid(NDim, sd, len(sd), Simil.NTheta(), nd, nn, ntn, arr, len(devs), pr, pc, Block,
                                    C.byref(self._g))
        else:
            idb = (C.c_ubyte * _lib.GRID_ID_BYTES)(*(UniqueId or bytes(_lib.GRID_ID_BYTES)))
            st = L.gogp_grid_create_rank(NDim, sd, len(sd), Simil.NTheta(), nd, nn, ntn, Device, Rank, World, pr, pc,
                                         Block, idb, C.byref(self._g))
        if st != _lib.OK:
            msg = self._err(st)
            self.close()
            raise GoGPPanic(st, msg)

    def _err(self, st):
        L = _lib.lib()
        m = L.gogp_grid_last_error(self._g).decode() if self._g else ""
        return m or L.gogp_status_string(st).decode()

    def close(self):
        if self._g:
            _lib.lib().gogp_grid_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _nts(self):
        return self.Simil.NTheta()

    def _ntn(self):
        return self.Noise.NTheta() if self.Noise is not None else 0

    def _upload(self):
        """gp.X / gp.Y are fields (tutorial/tutorial.go:114-115): upload them when they changed."""
        X = _flat(self.X, self.NDim)
        Y = _flat(self.Y, 0)
        if X.size != len(Y) * self.NDim:
            return _lib.BAD_ARGUMENT, "len(x) != len(y)"
        key = (id(self.X), id(self.Y), len(Y))
        if key != self._data_key:
            st = _lib.lib().gogp_grid_set_data(self._g, _lib.dptr(X), _lib.dptr(Y), len(Y))
            if st != _lib.OK:
                return st, self._err(st)
            self._data_key = key
        return _lib.OK, ""

    def Observe(self, x):
        """gp.GP.Observe, hyper-parameters only: x = [log theta_simil | log theta_noise]; panics like the
        reference on a bad length or a covariance that is not positive definite (gp/gp.go:398-405)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if len(x) != self._nts() + self._ntn():
            raise GoGPPanic(_lib.BAD_ARGUMENT, "len(x)")
        st, msg = self._upload()
        if st != _lib.OK:
            raise GoGPPanic(st, msg)
        lml = C.c_double()
        st = _lib.lib().gogp_grid_observe(self._g, _lib.dptr(x), C.byref(lml))
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        self.ThetaSimil = list(np.exp(x[:self._nts()]))
        self.ThetaNoise = list(np.exp(x[self._nts():]))
        return lml.value

    def Gradient(self):
        g = np.zeros(self._nts() + self._ntn())
        st = _lib.lib().gogp_grid_gradient(self._g, _lib.dptr(g), len(g))
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return g

    def Absorb(self, x, y):
        """Returns None or a GoGPError (the reference returns ``err``, gp/gp.go:80-87)."""
        self.X, self.Y = x, y
        if len(self.ThetaSimil) == 0:
            self.ThetaSimil = [0.0] * self._nts()
        if len(self.ThetaNoise) == 0:
            self.ThetaNoise = [0.0] * self._ntn()
        st, msg = self._upload()
        if st != _lib.OK:
            return GoGPError(st, msg)
        ts = np.array(self.ThetaSimil if self._nts() else [0.0], dtype=np.float64)
        tn = np.array(self.ThetaNoise if self._ntn() else [0.0], dtype=np.float64)
        st = _lib.lib().gogp_grid_absorb(self._g, _lib.dptr(ts), _lib.dptr(tn))
        return None if st == _lib.OK else GoGPError(st, self._err(st))

    def LML(self):
        v = C.c_double()
        st = _lib.lib().gogp_grid_lml(self._g, C.byref(v))
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return v.value

    def Alpha(self):
        n = len(_flat(self.Y, 0))
        a = np.zeros(n)
        st = _lib.lib().gogp_grid_get_alpha(self._g, _lib.dptr(a), n)
        if st != _lib.OK:
            raise GoGPPanic(st, self._err(st))
        return a

    def PhaseTimes(self):
        ms, cm = np.zeros(len(_lib.GRID_PHASES)), np.zeros(len(_lib.GRID_PHASES))
        _lib.lib().gogp_grid_phase_times(self._g, _lib.dptr(ms), _lib.dptr(cm))
        return dict(zip(_lib.GRID_PHASES, ms)), dict(zip(_lib.GRID_PHASES, cm))

    def Stats(self):
        s = np.zeros(12)
        _lib.lib().gogp_grid_stats(self._g, _lib.dptr(s))
        return {"peer_copy_bytes_received": int(s[8]), "collective_bytes_received": int(s[0]), "launches": int(s[1]), "pr": int(s[2]), "pc": int(s[3]),
                "block": int(s[4]), "device_bytes": int(s[5]), "nccl_version": int(s[6]), "world": int(s[7]), "eval_ms": float(s[9])}
