"""Batched orbit container: B independent KS orbits of one class and discretization, resident on one GPU.

Mirrors, with a leading batch axis, the part of the reference's ``Orbit``/``OrbitKS`` protocol that the solver
drivers touch (reference orbithunter/core.py: cost :986, dot :864, cdof :1617, constants :1645, increment :1760;
orbithunter/ks/orbits.py: transform :231, eqn :526, matvec :659, rmatvec :711, costgrad :757, precondition :1034,
from_numpy_array :305).  PyTorch owns the device memory; every numerical method is one call into libks_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

CLASS_IDS = {"OrbitKS": 0, "RelativeOrbitKS": 1, "ShiftReflectionOrbitKS": 2, "AntisymmetricOrbitKS": 3}
CLASS_NAMES = {v: k for k, v in CLASS_IDS.items()}
BASIS_IDS = {"field": 0, "spatial_modes": 1, "modes": 2}
PARAMETER_LABELS = ("t", "x", "s")


def default_constraints(cls_id):
    """ks/orbits.py:1731-1736; RelativeOrbitKS :3035-3036."""
    return {"t": False, "x": False, "s": cls_id != 1}


def constraint_mask(constraints):
    return sum(1 << i for i, lab in enumerate(PARAMETER_LABELS) if constraints.get(lab, True))


def basis_shape(cls_id, n, m, basis):
    """ks/orbits.py:1403-1412, :3254-3263, :3677-3686."""
    if basis == "field":
        return n, m
    if basis == "spatial_modes":
        return n, m - 2
    return (n - 1, m - 2) if cls_id in (0, 1) else (n - 1, m // 2 - 1)


_WORKSPACES = {}


def workspace(device, nbytes, stream=None):
    """Scratch buffer per (device, current stream), grown on demand (sized by *_workspace_bytes); never shrinks."""
    dev = device if isinstance(device, torch.device) else torch.device(device)
    if stream is None:
        stream = torch.cuda.current_stream(dev).cuda_stream
    key = (dev.index or 0, stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=dev)
        _WORKSPACES[key] = ws
    return ws


def _ptr(t):
    return None if t is None else t.data_ptr()


class BatchedOrbitKS:
    """B orbits: ``state`` (B, rows, cols) float64 CUDA tensor in ``basis``; ``parameters`` (B, 3) = (T, L, S)."""

    def __init__(self, state, parameters, cls="OrbitKS", basis="modes", discretization=None, constraints=None):
        self.cls_id = CLASS_IDS[cls] if isinstance(cls, str) else int(cls)
        if not torch.is_tensor(state):
            state = torch.as_tensor(np.ascontiguousarray(state, dtype=np.float64))
        if not torch.is_tensor(parameters):
            parameters = torch.as_tensor(np.ascontiguousarray(parameters, dtype=np.float64))
        if state.dim() != 3:
            raise ValueError('"state" must have shape (batch, rows, cols)')
        if basis not in BASIS_IDS:
            raise ValueError('basis not recognized; must equal "field" or "spatial_modes", or "modes"')
        if not state.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.KsB200Error("orbithunter_b200 needs a CUDA device; there is no CPU path")
            state = state.cuda()
        self.state = state.to(torch.float64).contiguous()
        self.parameters = parameters.to(self.state.device, torch.float64).reshape(-1, 3).contiguous()
        if self.parameters.shape[0] != self.state.shape[0]:
            raise ValueError("parameters must have shape (batch, 3)")
        self.basis = basis
        if discretization is None:
            r, c = self.state.shape[1:]
            if basis == "field":
                discretization = (r, c)
            elif basis == "spatial_modes":
                discretization = (r, c + 2)
            else:
                discretization = (r + 1, c + 2) if self.cls_id in (0, 1) else (r + 1, 2 * c + 2)
        self.discretization = tuple(int(x) for x in discretization)
        if tuple(self.state.shape[1:]) != basis_shape(self.cls_id, *self.discretization, basis):
            raise ValueError(f"state shape {tuple(self.state.shape[1:])} inconsistent with {self.discretization} in {basis}")
        cons = default_constraints(self.cls_id)
        if constraints:
            cons.update(constraints)
        self.constraints = cons

    @classmethod
    def from_seeds(cls, seeds, parameters, cls_name="OrbitKS", discretization=(32, 32), constraints=None, device=None, **kwargs):
        """Random initial conditions generated ON THE DEVICE: the batch equivalent of the reference's
        ``Cls(parameters=(T, L, S)).populate(attr="state", seed=s, discretization=(n, m), **kwargs)`` (ks/orbits.py:1738-1849)
        for every integer seed ``s`` in ``seeds`` (NumPy's legacy MT19937 / randn stream).  ``kwargs``: the reference's
        ``spatial_modulation``, ``temporal_modulation``, ``xshift`` (by name), ``tscale``, ``xscale``, ``xvar``, ``tvar``.
        ``parameters`` is (3,) or (B, 3).  RelativeOrbitKS: S is reset to 0 as the reference's populate does (:2872-2883)."""
        from ._cabi import SPATIAL_MODULATION_IDS, TEMPORAL_MODULATION_IDS, XSHIFT_IDS, KsPopulateOptions

        xs = kwargs.get("xshift", "sqrt")
        if not isinstance(xs, str):
            raise NotImplementedError("from_seeds: an array-valued xshift is only provided by the single-orbit populate()")
        nan = float("nan")
        opt = KsPopulateOptions(SPATIAL_MODULATION_IDS.get(kwargs.get("spatial_modulation", "laplace"), 0),
                                TEMPORAL_MODULATION_IDS.get(kwargs.get("temporal_modulation", "gaussian"), 0), XSHIFT_IDS.get(xs, 0), 0,
                                *(float(kwargs[k]) if kwargs.get(k) is not None else nan for k in ("tscale", "xscale", "xvar", "tvar")))
        if not torch.cuda.is_available():
            raise _lib.KsB200Error("orbithunter_b200 needs a CUDA device; there is no CPU path")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        cid = CLASS_IDS[cls_name] if isinstance(cls_name, str) else int(cls_name)
        seeds_np = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64).ravel())
        if seeds_np.size and (seeds_np.min() < 0 or seeds_np.max() > 2**32 - 1):
            raise ValueError("Seed must be between 0 and 2**32 - 1")
        B = int(seeds_np.size)
        n, m = (int(x) for x in discretization)
        par = torch.as_tensor(np.ascontiguousarray(parameters, dtype=np.float64)).reshape(-1, 3)
        if par.shape[0] == 1 and B != 1:
            par = par.expand(B, 3)
        par = par.contiguous().to(dev)
        if par.shape[0] != B:
            raise ValueError("parameters must have shape (3,) or (len(seeds), 3)")
        if cid == 1:
            par[:, 2] = 0.0
        sd = torch.as_tensor(seeds_np.astype(np.uint32).view(np.int32)).to(dev)
        state = torch.empty((B,) + basis_shape(cid, n, m, "modes"), dtype=torch.float64, device=dev)
        if B:
            with torch.cuda.device(dev):
                import ctypes

                _lib.check(_lib.lib().ks_populate_state_ex(cid, n, m, B, sd.data_ptr(), par.data_ptr(), ctypes.addressof(opt),
                                                           state.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return cls(state, par, cid, "modes", (n, m), constraints)

    # ---- bookkeeping -------------------------------------------------------------------------------------------
    @property
    def n(self):
        return self.discretization[0]

    @property
    def m(self):
        return self.discretization[1]

    @property
    def batch(self):
        return self.state.shape[0]

    @property
    def cmask(self):
        return constraint_mask(self.constraints)

    @property
    def cls_name(self):
        return CLASS_NAMES[self.cls_id]

    def __len__(self):
        return self.batch

    def _like(self, state=None, parameters=None, basis=None):
        """Another batch of the same class / grid / constraints around freshly computed tensors (no re-validation: the
        callers pass float64 contiguous CUDA tensors of the right shapes)."""
        out = object.__new__(BatchedOrbitKS)
        out.cls_id = self.cls_id
        out.state = self.state if state is None else state
        out.parameters = self.parameters if parameters is None else parameters
        out.basis = self.basis if basis is None else basis
        out.discretization = self.discretization
        out.constraints = dict(self.constraints)
        return out

    def copy(self):
        return self._like(self.state.clone(), self.parameters.clone())

    def _call(self, fn, *args):
        L = _lib.lib()
        dev = self.state.device
        nbytes = L.ks_workspace_bytes(self.cls_id, self.n, self.m, self.batch)
        if torch.cuda.current_device() == dev.index:  # the common case: no device switch, one stream lookup
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = workspace(dev, nbytes, stream)
            _lib.check(getattr(L, fn)(self.cls_id, self.n, self.m, self.batch, *args, ws.data_ptr(), ws.numel(), stream))
            return
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = workspace(dev, nbytes, stream)
            _lib.check(getattr(L, fn)(self.cls_id, self.n, self.m, self.batch, *args, ws.data_ptr(), ws.numel(), stream))

    def _need_modes(self):
        assert self.basis == "modes", "Convert to spatiotemporal Fourier mode basis first (reference ks/orbits.py:541)"

    # ---- reference protocol, batched ---------------------------------------------------------------------------
    def transform(self, to=None):
        if to not in BASIS_IDS:
            raise ValueError("Trying to transform to unrecognized basis.")
        if to == self.basis:
            return self
        out = torch.empty((self.batch,) + basis_shape(self.cls_id, self.n, self.m, to), dtype=torch.float64,
                          device=self.state.device)
        self._call("ks_transform", BASIS_IDS[self.basis], BASIS_IDS[to], _ptr(self.state), _ptr(out))
        return self._like(out, basis=to)

    def eqn(self, return_cost=False):
        self._need_modes()
        F = torch.empty_like(self.state)
        cost = torch.empty(self.batch, dtype=torch.float64, device=self.state.device)
        self._call("ks_eqn", _ptr(self.state), _ptr(self.parameters), _ptr(F), _ptr(cost))
        out = self._like(F)
        return (out, cost) if return_cost else out

    def cost(self, evaleqn=True):
        if evaleqn:
            return self.eqn(return_cost=True)[1]
        v = self.state.reshape(self.batch, -1)
        return 0.5 * (v * v).sum(dim=1)

    def matvec(self, other):
        self._need_modes()
        out = torch.empty_like(self.state)
        self._call("ks_matvec", _ptr(self.state), _ptr(self.parameters), _ptr(other.state), _ptr(other.parameters), self.cmask,
                   _ptr(out))
        return self._like(out)

    def jacobian(self):
        """Dense Jacobians of the whole batch (OrbitKS.jacobian, ks/orbits.py:557-657), as a [batch, Nm, cdof] tensor (a
        transposed view of the column-major array ks_jacobian fills)."""
        self._need_modes()
        L = _lib.lib()
        dev = self.state.device
        nm = self.state.shape[1] * self.state.shape[2]
        ncol = nm + len(self.free_indices())
        out = torch.empty((self.batch, ncol, nm), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = workspace(dev, L.ks_jacobian_workspace_bytes(self.cls_id, self.n, self.m, self.batch, self.cmask), stream)
            _lib.check(L.ks_jacobian(self.cls_id, self.n, self.m, self.batch, _ptr(self.state), _ptr(self.parameters), self.cmask,
                                     _ptr(out), ws.data_ptr(), ws.numel(), stream))
        return out.transpose(1, 2)

    def rmatvec(self, other):
        self._need_modes()
        out = torch.empty_like(self.state)
        op = torch.empty_like(self.parameters)
        self._call("ks_rmatvec", _ptr(self.state), _ptr(self.parameters), _ptr(other.state), self.cmask, _ptr(out), _ptr(op))
        return self._like(out, op)

    def costgrad(self, preconditioning=False, return_eqn=False):
        """(cost, gradient orbit[, F orbit]): eqn + costgrad of the reference fused into one launch."""
        self._need_modes()
        g = torch.empty_like(self.state)
        gp = torch.empty_like(self.parameters)
        cost = torch.empty(self.batch, dtype=torch.float64, device=self.state.device)
        F = torch.empty_like(self.state) if return_eqn else None
        self._call("ks_costgrad", _ptr(self.state), _ptr(self.parameters), self.cmask, int(bool(preconditioning)), _ptr(cost),
                   _ptr(g), _ptr(gp), _ptr(F))
        grad = self._like(g, gp)
        return (cost, grad, self._like(F)) if return_eqn else (cost, grad)

    def precondition(self, pmult=None, pexp=(1, 4)):
        """ks/orbits.py:1034-1089: rescale this (gradient / correction) orbit with respect to ``pmult``'s (T, L)."""
        ref = self.parameters if pmult is None else pmult
        out = torch.empty_like(self.state)
        gp = torch.empty_like(self.parameters)
        L = _lib.lib()
        with torch.cuda.device(self.state.device):
            _lib.check(L.ks_precondition(self.cls_id, self.n, self.m, self.batch, _ptr(self.state), _ptr(self.parameters),
                                         _ptr(ref.contiguous()), self.cmask, float(pexp[0]), float(pexp[1]), _ptr(out), _ptr(gp),
                                         torch.cuda.current_stream().cuda_stream))
        return self._like(out, gp)

    def _deriv(self, axis, order):
        out = torch.empty_like(self.state)
        L = _lib.lib()
        with torch.cuda.device(self.state.device):
            _lib.check(L.ks_deriv(self.cls_id, self.n, self.m, self.batch, axis, int(order), self.state.shape[1],
                                  self.state.shape[2], _ptr(self.state), _ptr(self.parameters), _ptr(out),
                                  torch.cuda.current_stream().cuda_stream))
        return self._like(out)

    def dt(self, order=1, return_basis=None):
        """ks/orbits.py:379-440: spectral time derivative, computed in the modes basis."""
        back = self.basis if return_basis is None else return_basis
        return self.transform("modes")._deriv(0, order).transform(back)

    def dx(self, order=1, computation_basis=None, return_basis=None):
        """ks/orbits.py:442-524 and the SR/Anti overrides :3190-3227, :3583-3619 (odd orders in spatial_modes)."""
        back = self.basis if return_basis is None else return_basis
        half = self.cls_id in (2, 3)
        if computation_basis is None:
            computation_basis = "spatial_modes" if (half and order % 2) else "modes"
        if computation_basis not in ("modes", "spatial_modes"):
            raise ValueError(f"dx(computation_basis={computation_basis}); invalid basis for spectral differentiation. ")
        return self.transform(computation_basis)._deriv(1, order).transform(back)

    def resize(self, *new_discretization):
        """Orbit.resize for the batch (core.py:1167-1226; _pad / _truncate ks/orbits.py:1428-1564 and the SR / Anti overrides):
        zero padding / truncation of the spatiotemporal modes on the device; comes back in the basis of self."""
        new = new_discretization[0] if len(new_discretization) == 1 and not np.isscalar(new_discretization[0]) else new_discretization
        n2, m2 = (int(x) for x in new)
        if (n2, m2) == self.discretization:
            return self
        if n2 % 2 or m2 % 2:
            raise ValueError("New discretization size must be an even number, preferably a power of 2")
        src = self.transform("modes")
        out = torch.empty((self.batch,) + basis_shape(self.cls_id, n2, m2, "modes"), dtype=torch.float64, device=self.state.device)
        with torch.cuda.device(self.state.device):
            _lib.check(_lib.lib().ks_resize_modes(self.cls_id, self.n, self.m, n2, m2, self.batch, _ptr(src.state), _ptr(out),
                                                  torch.cuda.current_stream().cuda_stream))
        res = self._like(out, basis="modes")
        res.discretization = (n2, m2)
        return res.transform(self.basis)

    def rescale(self, magnitude, method="inf"):
        """Orbit.rescale (core.py:1831-1853) for every orbit of the batch: field norm ('inf', 'L1', 'L2') set to ``magnitude``,
        or the signed power ('LP'); result in the basis of self."""
        mid = {"inf": 0, "L1": 1, "L2": 2, "LP": 3}.get(method)
        if mid is None:
            raise ValueError("Unrecognizable method.")
        f = self.transform("field")
        f = f._like(f.state.clone()) if f is self else f
        with torch.cuda.device(self.state.device):
            _lib.check(_lib.lib().ks_rescale_field(self.n, self.m, self.batch, _ptr(f.state), float(magnitude), mid, None,
                                                   torch.cuda.current_stream().cuda_stream))
        return f.transform(self.basis)

    def preprocess_codes(self):
        """Per-orbit classification of OrbitKS.preprocess (ks/orbits.py:1566-1596; Relative :2969-3008): (code, flip) with
        code 0 = keep, 1 = zero solution, 2 = equilibrium (no temporal variation); flip (RelativeOrbitKS only): the orbit with
        the shift negated has the smaller cost and is the one the reference carries on with."""
        mo = self.transform("modes")
        code = torch.empty(self.batch, dtype=torch.int32, device=self.state.device)
        flip = torch.zeros(self.batch, dtype=torch.bool, device=self.state.device)
        if self.cls_id == 1:
            mirrored = mo._like(parameters=mo.parameters * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64, device=self.state.device))
            flip = mo.cost() > mirrored.cost()
        with torch.cuda.device(self.state.device):
            _lib.check(_lib.lib().ks_preprocess(self.cls_id, self.n, self.m, self.batch, _ptr(mo.state), _ptr(mo.parameters),
                                                _ptr(code), None, torch.cuda.current_stream().cuda_stream))
        return code.cpu().numpy(), flip.cpu().numpy()

    def dot(self, other):
        return (self.state.reshape(self.batch, -1) * other.state.reshape(self.batch, -1)).sum(dim=1)

    def norm(self):
        return torch.linalg.vector_norm(self.state.reshape(self.batch, -1), dim=1)

    def increment(self, other, step_size=1.0):
        """core.py:1760-1796; ``step_size`` may be a scalar or a (B,) tensor."""
        if torch.is_tensor(step_size):
            s3, s2 = step_size.reshape(-1, 1, 1), step_size.reshape(-1, 1)
        else:
            s3 = s2 = step_size
        return self._like(self.state + s3 * other.state, self.parameters + s2 * other.parameters)

    def free_indices(self):
        return [i for i, lab in enumerate(PARAMETER_LABELS) if not self.constraints.get(lab, True)]

    def cdof(self):
        """[state.ravel(), free parameters] per orbit; core.py:1617-1643."""
        return torch.cat((self.state.reshape(self.batch, -1), self.parameters[:, self.free_indices()]), dim=1)

    def from_cdof(self, vec, constants=None):
        """Inverse of cdof; constrained parameters from ``constants`` (B,3) or 0 (ks/orbits.py:343-377)."""
        size = self.state.shape[1] * self.state.shape[2]
        params = torch.zeros_like(self.parameters) if constants is None else constants.clone()
        if constants is not None:
            params[:, self.free_indices()] = 0.0
        for slot, i in enumerate(self.free_indices()):
            params[:, i] = vec[:, size + slot]
        return self._like(vec[:, :size].reshape(self.state.shape).contiguous(), params)

    # ---- host <-> device ---------------------------------------------------------------------------------------
    def cpu_state(self):
        return self.state.cpu().numpy()

    def cpu_parameters(self):
        return self.parameters.cpu().numpy()

    def __getitem__(self, i):
        """Single-orbit view as a drop-in OrbitKS-family instance (host arrays)."""
        from .ks import CLASSES

        p = tuple(float(x) for x in self.parameters[i].cpu().numpy())
        return CLASSES[self.cls_name](state=self.state[i].cpu().numpy(), basis=self.basis, parameters=p,
                                      discretization=self.discretization, constraints=dict(self.constraints))
