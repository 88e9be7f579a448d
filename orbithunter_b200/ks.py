"""Single-orbit drop-in classes: OrbitKS, RelativeOrbitKS, ShiftReflectionOrbitKS, AntisymmetricOrbitKS.

Same constructor, attribute and method names as the reference's equation plugin (orbithunter/ks/orbits.py:24,
:2554, :3582, :3189 and the generic container orbithunter/core.py:10) for the hot path named by the north star:
``transform, eqn, matvec, rmatvec, cost, costgrad, precondition, cdof, constants, from_numpy_array, increment,
dot, norm, jacobian, populate``.  State and parameters are host objects (NumPy array, tuple of floats) exactly as in
the reference; every numerical method uploads a batch of one to the GPU, runs the same libks_b200.so kernels the
batched container uses, and returns a new instance.  There is no NumPy implementation behind these methods.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .batch import BatchedOrbitKS, basis_shape

_TWO_PI = 2 * np.pi


class OrbitKS:
    _cls_name = "OrbitKS"

    # ---- static description (reference ks/orbits.py:138-229, :1717-1736) ---------------------------------------
    @staticmethod
    def bases_labels():
        return "field", "spatial_modes", "modes"

    @staticmethod
    def parameter_labels():
        return "t", "x", "s"

    @staticmethod
    def discretization_labels():
        return "n", "m"

    @staticmethod
    def dimension_labels():
        return "t", "x"

    @staticmethod
    def minimal_shape():
        return 2, 4

    @staticmethod
    def minimal_shape_increments():
        return 2, 2

    @staticmethod
    def _default_shape():
        return 32, 32

    @classmethod
    def _default_parameter_ranges(cls):
        return {"t": (20, 200), "x": (20, 100)}

    @classmethod
    def _default_constraints(cls):
        return {"t": False, "x": False, "s": True}

    def periodic_dimensions(self):
        return True, True

    # ---- construction (core.py:111-134, ks/orbits.py:100-120, :1851-1897) ----------------------------------------
    def __init__(self, state=None, basis=None, parameters=None, discretization=None, constraints=None, **kwargs):
        if state is None or basis is None or discretization is None:
            self._parse_state(state, basis)
        else:
            self.state, self.basis, self.discretization = state, basis, tuple(discretization)
        if parameters is None or constraints is None:
            self._parse_parameters(parameters, constraints)
        else:
            self.parameters, self.constraints = parameters, constraints
        if self.parameters:
            self.parameters = tuple(float(p) for p in self.parameters)
        self._workers = kwargs.get("_workers", 1)

    def _mode_discretization(self, shape):
        return shape[0] + 1, shape[1] + 2

    def _parse_state(self, state, basis):
        if isinstance(state, np.ndarray):
            if state.ndim != 2:
                raise ValueError('"state" array must be two-dimensional')
            self.state = state
        else:
            self.state = np.array([], dtype=float).reshape(0, 0)
        if self.state.size > 0:
            if basis is None:
                raise ValueError("basis must be provided when state is provided")
            if basis == "modes":
                n, m = self._mode_discretization(self.state.shape)
            elif basis == "field":
                n, m = self.state.shape
            elif basis == "spatial_modes":
                n, m = self.state.shape[0], self.state.shape[1] + 2
            else:
                raise ValueError('basis not recognized; must equal "field" or "spatial_modes", or "modes"')
            if n < self.minimal_shape()[0] or m < self.minimal_shape()[1]:
                warnings.warn("\nminimum discretization requirements not met; methods may not work as intended.", RuntimeWarning)
            self.basis, self.discretization = basis, (n, m)
        else:
            self.basis, self.discretization = None, None

    def _parse_parameters(self, parameters, constraints):
        defaults = self._default_constraints()
        given = defaults if constraints is None else constraints
        self.constraints = {k: (given.get(k, True) if k in defaults else True) for k in self.parameter_labels()}
        if parameters is None:
            self.parameters = None
        elif isinstance(parameters, tuple):
            labels = self.parameter_labels()
            vals = list(parameters)[: len(labels)]
            self.parameters = tuple(vals + [0] * (len(labels) - len(vals)))
        else:
            raise TypeError('"parameters" is required to be a tuple or NoneType. singleton parameters need to be cast as tuple (val,).')

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if attr in ("t", "x", "s"):
            params = self.__dict__.get("parameters")
            if params is None:
                raise AttributeError(f"Cannot retrieve '{attr}' when parameters attribute is NoneType")
            return params["txs".index(attr)]
        if attr in ("n", "m"):
            disc = self.__dict__.get("discretization")
            if disc is None:
                raise AttributeError(f"Cannot retrieve '{attr}' when discretization attribute is NoneType")
            return disc["nm".index(attr)]
        raise AttributeError(f"{self.__class__.__name__} has no attribute {attr}")

    def __str__(self):
        return self.__class__.__name__

    def __repr__(self):
        p = None if self.parameters is None else tuple(round(float(x), 3) for x in self.parameters)
        return f'{self.__class__.__name__}({{"shape": {list(self.shape)}, "basis": "{self.basis}", "(t, x, s)": {p}}})'

    @property
    def shape(self):
        return self.state.shape

    @property
    def size(self):
        return self.state.size

    @property
    def ndim(self):
        return self.state.ndim

    def shapes(self):
        n, m = self.discretization
        return tuple(basis_shape(self._cls_id(), n, m, b) for b in self.bases_labels())

    def dimensions(self):
        return self.t, self.x

    def copy(self):
        return self.__class__(**{k: (v.copy() if hasattr(v, "copy") else v) for k, v in vars(self).items()})

    def _new(self, **changes):
        return self.__class__(**{**vars(self), **changes})

    @classmethod
    def _cls_id(cls):
        from .batch import CLASS_IDS

        return CLASS_IDS[cls._cls_name]

    # ---- device round trip ---------------------------------------------------------------------------------------
    def _batched(self, state=None, basis=None, parameters=None):
        st = self.state if state is None else state
        pr = self.parameters if parameters is None else parameters
        return BatchedOrbitKS(torch.as_tensor(np.ascontiguousarray(st, dtype=np.float64)[None]),
                              torch.as_tensor(np.asarray(pr, dtype=np.float64)[None]), self._cls_id(),
                              self.basis if basis is None else basis, self.discretization, self.constraints)

    @staticmethod
    def _host(t):
        return t[0].cpu().numpy()

    # ---- arithmetic (core.py:136-451) ----------------------------------------------------------------------------
    def _other(self, other):
        return other.state if isinstance(other, OrbitKS) else other

    def __add__(self, other):
        return self._new(state=self.state + self._other(other))

    __radd__ = __add__

    def __sub__(self, other):
        return self._new(state=self.state - self._other(other))

    def __rsub__(self, other):
        return self._new(state=self._other(other) - self.state)

    def __mul__(self, other):
        return self._new(state=np.multiply(self.state, self._other(other)))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self._new(state=np.divide(self.state, self._other(other)))

    def __neg__(self):
        return self._new(state=-self.state)

    def dot(self, other):
        return float(np.dot(self.state.ravel(), other.state.ravel()))

    def norm(self, order=None):
        return np.linalg.norm(self.state.ravel(), ord=order)

    # ---- transforms (ks/orbits.py:231-303) -----------------------------------------------------------------------
    def transform(self, to=None, array=False, inplace=False):
        if self.state is None or self.state.size == 0:
            raise ValueError(f"Trying to transform an unpopulated {self} instance.")
        if self.basis is None:
            raise ValueError("Trying to transform state with unknown basis")
        if to not in self.bases_labels():
            raise ValueError("Trying to transform to unrecognized basis.")
        if to == self.basis:
            return self.state if array else self
        out = self._host(self._batched().transform(to).state)
        if inplace:
            self.state, self.basis = out, to
            return self.state if array else self
        return out if array else self._new(state=out, basis=to)

    # ---- spectral derivatives (ks/orbits.py:379-524; SR/Anti :3190-3227, :3583-3619) ------------------------------------
    def dt(self, order=1, array=False, **kwargs):
        if getattr(self, "frame", "comoving") != "comoving":
            raise ValueError(f"Attempting to compute time derivative of {self} in physical reference frame.")
        if kwargs.get("inplace", False):  # only the in-place form honours return_basis (:404-420)
            back = kwargs.get("return_basis", self.basis)
            self.state = self._host(self._batched().dt(order, return_basis=back).state)
            self.basis = back
            return self.state if array else self
        if array:  # the reference returns the modes-basis array in this case (:432-434)
            return self._host(self._batched().dt(order, return_basis="modes").state)
        # not in place: the derivative comes back in the basis of self, whatever return_basis says (:435-440)
        return self._new(state=self._host(self._batched().dt(order, return_basis=self.basis).state), basis=self.basis)

    def dx(self, **kwargs):
        order, array = kwargs.get("order", 1), kwargs.get("array", False)
        half = self._cls_id() in (2, 3)
        comp = kwargs.get("computation_basis", "modes")
        if half and order % 2:
            comp = "spatial_modes"  # odd orders of the discrete-symmetry classes NEED the spatial mode basis (:3222-3224, :3614-3616)
        if comp not in ("modes", "spatial_modes"):
            raise ValueError(f"{str(self)}.dx(computation_basis={comp}); invalid basis for spectral differentiation. ")
        back = kwargs.get("return_basis", self.basis)
        if kwargs.get("inplace", False):
            self.state = self._host(self._batched().dx(order, comp, return_basis=back).state)
            self.basis = back
            return self.state if array else self
        if array:  # array=True returns the derivative in the computation basis (:518-519)
            return self._host(self._batched().dx(order, comp, return_basis=comp).state)
        return self._new(state=self._host(self._batched().dx(order, comp, return_basis=back).state), basis=back)

    # ---- governing equation and its linearisations ------------------------------------------------------------------
    def eqn(self, **kwargs):
        assert self.basis == "modes", "Convert to spatiotemporal Fourier mode basis before computing K-S equation DAEs."
        return self._new(state=self._host(self._batched().eqn().state))

    def cost(self, evaleqn=True):
        if evaleqn:
            return float(self.transform(to="modes")._batched().cost()[0].item())
        v = self.state.ravel()
        return 0.5 * float(v.dot(v))

    def matvec(self, other, **kwargs):
        assert (self.basis == "modes") and (other.basis == "modes")
        vb = self._batched(state=other.state, parameters=other.parameters)
        return self._new(state=self._host(self._batched().matvec(vb).state))

    def rmatvec(self, other, **kwargs):
        assert (self.basis == "modes") and (other.basis == "modes")
        r = self._batched().rmatvec(self._batched(state=other.state))
        return self._new(state=self._host(r.state), parameters=tuple(float(x) for x in self._host(r.parameters)))

    def costgrad(self, *args, **kwargs):
        """J^T F (ks/orbits.py:757-791).  With no argument F is evaluated and the fused kernel is used."""
        pc = bool(kwargs.get("preconditioning", False))
        if args and args[0] is not None:
            grad = self.rmatvec(args[0])
            if pc:
                grad = grad.precondition(**{"pmult": self.preconditioning_parameters(), **kwargs})
            return grad
        _, g = self._batched().costgrad(preconditioning=pc)
        return self._new(state=self._host(g.state), parameters=tuple(float(x) for x in self._host(g.parameters)))

    def preconditioning_parameters(self):
        return (self.t, self.n), (self.x, self.m)

    def precondition(self, **kwargs):
        """Diagonal rescaling 1/(|w| + q^2 + q^4), t/T, x/L^4 (ks/orbits.py:1034-1089), on the device."""
        (T, n), (L, m) = kwargs.get("pmult", self.preconditioning_parameters())
        if (n, m) != tuple(self.discretization):
            raise ValueError("pmult discretization must match the orbit being preconditioned")
        ref = torch.tensor([[float(T), float(L), 0.0]], dtype=torch.float64, device="cuda")
        r = self._batched(basis="modes").precondition(ref, kwargs.get("pexp", (1, 4)))
        return self._new(state=self._host(r.state), basis="modes", parameters=tuple(float(x) for x in self._host(r.parameters)))

    def preconditioner(self, **kwargs):
        """Diagonal preconditioner as a SciPy LinearOperator (ks/orbits.py:1091-1141); the diagonal comes from the device."""
        from scipy.sparse.linalg import LinearOperator

        ones = self._new(state=np.ones(self.shape), parameters=(1.0, 1.0, 1.0), basis="modes")
        mult = ones.precondition(**{"pmult": self.preconditioning_parameters(), **kwargs})
        diag = mult.cdof().reshape(-1, 1)

        def mv(v):
            return v * diag.reshape(v.shape)

        def mm(V):
            return diag.reshape(-1, 1) * V

        return LinearOperator(shape=(diag.size, diag.size), matvec=mv, rmatvec=mv, matmat=mm, rmatmat=mm)

    def jacobian(self, **kwargs):
        """Dense (Nm x cdof) Jacobian (ks/orbits.py:557-657): every column is one matvec against a unit vector, all in one
        launch on the GPU (ks_jacobian)."""
        assert self.basis == "modes", "Convert to spatiotemporal Fourier mode basis before computing Jacobian"
        return self._host(self._batched(basis="modes").jacobian().contiguous())

    # ---- vector packing (core.py:1617-1664, :1760-1796; ks/orbits.py:305-377) ----------------------------------------
    def cdof(self):
        free = tuple(p for p, lab in zip(self.parameters or (), self.parameter_labels()) if not self.constraints[lab])
        return np.concatenate((self.state.ravel(), free), axis=0).reshape(-1, 1)

    def constants(self):
        return tuple(p for p, lab in zip(self.parameters or (), self.parameter_labels()) if self.constraints[lab])

    def orbit_vector(self):
        return np.concatenate((self.state.ravel(), tuple(self.parameters or ())), axis=0).reshape(-1, 1)

    def from_numpy_array(self, cdof, *args, **kwargs):
        flat = np.asarray(cdof).ravel()
        free = [float(p) for p in flat[self.size:].real]
        const = list(args)
        params = None
        if self.parameters is not None:
            vals = []
            for lab in self.parameter_labels():
                if not self.constraints.get(lab, True) and free:
                    vals.append(free.pop(0))
                else:
                    vals.append(float(const.pop(0)) if const else 0.0)
            params = tuple(vals)
        return self._new(state=np.reshape(flat[: self.size], self.shape), parameters=params, **kwargs)

    def increment(self, other, step_size=1, **kwargs):
        params = None
        if self.parameters is not None:
            params = tuple(a + step_size * b for a, b in zip(self.parameters, other.parameters))
        return self.__class__(**{**vars(self), **kwargs, "state": self.state + step_size * other.state, "parameters": params})

    # ---- seeds (ks/orbits.py:1738-1849, core.py:2016-2044, :2434-2514; default modulations only) ----------------------
    # ---- rediscretization (reference core.py:1167-1226, ks/orbits.py:1428-1564): zero padding / truncation of the modes ----
    def resize(self, *new_discretization, **kwargs):
        new_shape = new_discretization
        if len(new_shape) == 1 and isinstance(new_shape[0], tuple):
            new_shape = tuple(new_shape[0])
        if not new_shape:  # core.py:1200-1203: the class conventions decide
            new_shape = self.dimension_based_discretization(self.dimensions(), **kwargs)
        new_shape = tuple(int(x) for x in new_shape)
        out = self.copy().transform(to="modes")
        if tuple(self.discretization) != new_shape:
            for ax, (old, new, min_size) in enumerate(zip(self.discretization, new_shape, self.minimal_shape())):
                if new < min_size:
                    raise ValueError("minimum discretization requirements not met during resize.")
                if new < old:
                    out = out._truncate(new, axis=ax)
                elif new > old:
                    out = out._pad(new, axis=ax)
        return out.transform(to=self.basis)

    @staticmethod
    def _check_even(size):
        if np.mod(size, 2):
            raise ValueError("New discretization size must be an even number, preferably a power of 2")

    def _pad(self, size, axis=0, **kwargs):
        modes = self.transform(to="modes")
        self._check_even(size)
        n, m = modes.discretization
        if axis == 0:
            h, padding = n // 2 - 1, (size - n) // 2
            padded = np.concatenate((modes.state[:-h, :], np.pad(modes.state[-h:, :], ((padding, padding), (0, 0)))), axis=0)
            padded = padded * np.sqrt(size / n)
            disc = (size, m)
        else:
            padded = self._pad_space(modes.state, m, size) * np.sqrt(size / m)
            disc = (n, size)
        return self._new(state=padded, basis="modes", discretization=disc).transform(to=self.basis)

    def _truncate(self, size, axis=0):
        modes = self.transform(to="modes")
        self._check_even(size)
        n, m = modes.discretization
        keep = int(size // 2) - 1
        if axis == 0:
            h = n // 2 - 1
            first, second = modes.state[: keep + 1, :], modes.state[-h:, :][:keep, :]
            truncated = np.sqrt(size / n) * np.concatenate((first, second), axis=0)
            disc = (size, m)
        else:
            truncated = np.sqrt(size / m) * self._truncate_space(modes.state, m, keep)
            disc = (n, size)
        return self._new(state=truncated, basis="modes", discretization=disc).transform(to=self.basis)

    @staticmethod
    def _pad_space(state, m, size):  # columns [Re_x | Im_x]: zeros after each block (ks/orbits.py:1470-1480)
        k, padding = m // 2 - 1, (size - m) // 2
        return np.concatenate((state[:, :-k], np.pad(state[:, -k:], ((0, 0), (padding, padding)))), axis=1)

    @staticmethod
    def _truncate_space(state, m, keep):  # (ks/orbits.py:1541-1550)
        k = m // 2 - 1
        return np.concatenate((state[:, :keep], state[:, -k:][:, :keep]), axis=1)

    def populate(self, attr="all", **kwargs):
        if attr in ("all", "parameters"):
            self._populate_parameters(**kwargs)
        if attr in ("all", "state"):
            if self.size == 0 or kwargs.get("overwrite", False):
                self._populate_state(**kwargs)
            else:
                raise ValueError(" overwriting a non-empty state requires overwrite=True. ")
        return self

    def _populate_parameters(self, **kwargs):
        if isinstance(kwargs.get("seed", None), int):
            np.random.seed(kwargs["seed"])
        ranges = kwargs.get("parameter_ranges", self._default_parameter_ranges())
        current = list(self.parameters or [None] * 3)[:3] + [None] * (3 - len(self.parameters or [None] * 3))
        out = []
        for lab, val in zip(self.parameter_labels(), current):
            gen = ranges.get(lab, (0, 0))
            if kwargs.get("overwrite", False) or val is None:
                if isinstance(gen, tuple) and len(gen) == 2:
                    val = gen[0] + (gen[1] - gen[0]) * np.random.rand()
                elif np.isscalar(gen):
                    val = gen
                else:
                    val = gen[np.random.choice(range(len(gen)))]
            out.append(float(val))
        self.parameters = tuple(out)

    # modulation profiles of the seed generator (ks/orbits.py:1797-1842): name -> f(index table, centre, variance, orbit)
    _SPATIAL_MODULATIONS = {
        "gaussian": lambda k, c, var, o: np.exp(-((k - c) ** 2) / (2 * var)),
        "laplace": lambda k, c, var, o: np.exp(-1.0 * np.abs(k - c) / np.sqrt(var)),
        "laplace_sqrt": lambda k, c, var, o: np.exp(-1.0 * np.sqrt(np.abs(k - c)) / np.sqrt(var)),
        "flat_top": lambda k, c, var, o: np.exp(-((np.abs(k - c) / var) ** 5)),
    }
    _TEMPORAL_MODULATIONS = {
        "gaussian": lambda j, c, var, o: np.exp(-((j - c) ** 2) / (2 * var)),
        "laplace": lambda j, c, var, o: np.exp(-1.0 * np.abs(j - c) / np.sqrt(var)),
        "flat_top": lambda j, c, var, o: np.exp(-((np.abs(j - c) / var) ** 5)),
        "laplace_plateau": lambda j, c, var, o: np.where(j < c, 1.0, np.exp(-1.0 * np.abs(j - c) / np.sqrt(var))),
        "truncate": lambda j, c, var, o: np.where(j > c, 0, 1),
    }

    @classmethod
    def dimension_based_discretization(cls, dimensions, **kwargs):
        """Grid size the reference's conventions assign to (T, L) (ks/orbits.py:1638-1694): powers of two near T/2 and
        sqrt(2) L by default ('coarse' / 'fine' / 'power' variants), never below 32 (16 for 'coarse') nor the class minimum;
        ``discretization=(N, None)`` pins an axis."""
        resolution = kwargs.get("resolution", "default")
        forced = kwargs.get("discretization", None) or (None, None)
        # per axis: (exponent offset applied to log2 of the dimension, floor) for the power-of-two rules
        rules = {"coarse": ((-2, 16), (-1, 16)), "fine": ((1, 32), (2, 32)), "default": ((-1, 32), (0.5, 32))}
        out = []
        for axis, (dim, pinned) in enumerate(zip(tuple(dimensions)[:2], forced)):
            if pinned is not None:
                out.append(pinned)
                continue
            if dim == 0:
                size = cls._default_shape()[axis]
            elif resolution == "power":
                size = max(2 * (int(4 * dim ** 0.5) // 2), 32)
            else:
                shift, floor_ = rules.get(resolution, rules["default"])[axis]
                size = max(2 ** int(np.log2(dim) + shift), floor_)
            out.append(max(size, cls.minimal_shape()[axis]))
        return tuple(out)

    def _populate_state(self, **kwargs):
        """Random spatiotemporal Fourier modes (ks/orbits.py:1738-1849): legacy ``np.random.seed(seed); randn(rows, cols)``,
        multiplied by a spatial profile centred on xscale (+ a time-dependent shift ``xshift``) and a temporal profile
        centred on tscale, then renormalised to the norm the unmodulated draw had."""
        T, L = self.t, self.x
        tscale = kwargs.get("tscale", int(np.round(T / 20.0)))
        xscale = kwargs.get("xscale", int(np.round(L / (_TWO_PI * np.sqrt(2)))))
        xvar = kwargs.get("xvar", max([np.sqrt(xscale), 1]))
        tvar = kwargs.get("tvar", max([np.sqrt(tscale), 1]))
        np.random.seed(kwargs.get("seed", None))  # legacy global MT19937 stream, as in the reference
        n, m = self.dimension_based_discretization(self.dimensions(), **kwargs)
        if n < self.minimal_shape()[0] or m < self.minimal_shape()[1]:
            warnings.warn("\nminimum discretization requirements not met; methods may not work as intended.", RuntimeWarning)
        self.discretization = (int(n), int(m))
        rows, cols = self.shapes()[2]
        # the reference truncates its floating-point frequency tables (spatial_frequencies(2 pi, m, 1), :1788-1791) with
        # .astype(int); for m, n that are not powers of two a few entries land on kappa - 1 -- same expression here
        kx = np.abs((2 * np.pi * m / (2 * np.pi)) * (np.arange(0, m // 2 + 1) * (1.0 / m))[1:-1]).astype(int)
        space_ = np.concatenate((kx, kx))[None, :cols]
        jt = np.abs((-1 * (2 * np.pi * n / (2 * np.pi))) * (np.arange(0, n // 2 + 1) * (1.0 / n))[1:-1]).astype(int)
        time_ = np.concatenate(([0], jt, jt))[:, None]
        random_modes = np.random.randn(rows, cols)
        norm0 = np.linalg.norm(random_modes)
        xshift = kwargs.get("xshift", "sqrt")  # space-time coupling of the spatial centre (:1797-1805)
        if isinstance(xshift, np.ndarray):
            centre = xscale + xshift
        else:
            centre = xscale + {"sqrt": np.sqrt, "log": lambda j: np.log(1 + j), "lin": lambda j: j}.get(xshift, lambda j: 0 * j)(time_)
        smod = kwargs.get("spatial_modulation", "laplace")
        tmod = kwargs.get("temporal_modulation", "gaussian")
        if smod in ("plateau_linear", "exponential_linear"):  # profiles of the linear KS spectrum q^2 - q^4 (:1816-1824)
            q = _TWO_PI * space_ / L
            with np.errstate(divide="ignore"):
                growth = q ** 2 - q ** 4
                prof = (np.divide(1, growth) if smod == "plateau_linear" else np.exp(growth)) * np.ones(time_.shape)
            if smod == "plateau_linear":
                prof[np.broadcast_to(space_ <= centre, prof.shape)] = 1
        elif smod in self._SPATIAL_MODULATIONS:
            prof = self._SPATIAL_MODULATIONS[smod](space_, centre, xvar, self)
        else:
            prof = np.ones(random_modes.shape)
        modes = np.multiply(prof, random_modes)
        tprof = self._TEMPORAL_MODULATIONS[tmod](time_, tscale, tvar, self) if tmod in self._TEMPORAL_MODULATIONS else np.ones(random_modes.shape)
        modes = np.multiply(tprof, modes)
        self.state = (norm0 / np.linalg.norm(modes)) * modes
        self.basis = "modes"

    # ---- on-disk format (core.py:1855-2001, ks/orbits.py:1598-1636) -----------------------------------------------------
    def filename(self, extension=".h5", decimals=3, cls_name=True):
        """Conventional file name (core.py:1952-2001): class name + one '_<label><value with p for the point>' per non-zero
        dimension, e.g. OrbitKS_t44p000_x33p400.h5."""
        dims = "".join(f"_{lab}" + f"{d:.{decimals}f}".replace(".", "p") for lab, d in zip(self.dimension_labels(), self.dimensions())
                       if d != 0 and d is not None)
        if not cls_name and dims:
            return (dims + extension).lstrip("_")
        return (self.__class__.__name__ + dims + extension).lstrip("_")

    def to_h5(self, filename=None, dataname=None, h5mode="a", verbose=False, include_cost=True, **kwargs):
        """OrbitKS.to_h5 (ks/orbits.py:1598-1636: the cost is stored by default) through the package's own HDF5 writer."""
        from .h5io import to_h5

        return to_h5(self, filename=filename or "", groupname=kwargs.get("groupname", ""), dataname=dataname or "", h5mode=h5mode,
                     verbose=verbose, include_cost=include_cost)

    def rescale(self, magnitude, method="inf"):
        """Orbit.rescale (core.py:1831-1853): set the size of the FIELD -- max-norm ('inf'), 'L1' (induced matrix 1-norm as
        numpy.linalg.norm(ord=1) gives for a 2-d array), 'L2' (Frobenius), or the signed power |u|^magnitude ('LP') --
        and return in the basis of self.  The norm and the scaling run on the device (ks_rescale)."""
        if method not in ("inf", "L1", "L2", "LP"):
            raise ValueError("Unrecognizable method.")
        out = self._batched().rescale(float(magnitude), method)
        return self._new(state=self._host(out.state), basis=self.basis)

    def preprocess(self):
        """Equilibrium / zero-solution check of a converged orbit (ks/orbits.py:1566-1596; Relative :2969-3008).  The norms
        of u and u_t are taken on the device (ks_preprocess).  The reference converts such orbits to its Equilibrium
        classes, which are outside this package's scope: a vanishing field comes back as zeros of this class, a genuine
        equilibrium is returned unchanged with ``.equilibrium = True``; everything else is ``self``."""
        code, flip = self._batched().preprocess_codes()
        code, flip = int(code[0]), bool(flip[0])
        if code == 0:
            if flip:  # RelativeOrbitKS: the mirrored shift has the smaller cost (:2974-2983)
                return self._new(parameters=(self.t, self.x, -self.s))
            return self
        if code == 1:
            return self._new(state=np.zeros(self.shapes()[0]), basis="field").transform(to=self.basis)
        out = self.copy()
        out.equilibrium = True
        return out

    # ---- symmetry operations and joining (ks/orbits.py:1143-1232, :1355-1401; core.py:1282-1380) ------------------------
    def roll(self, shift, axis=0):
        """np.roll of the field along ``axis`` (ks/orbits.py:1172-1194), returned in the basis of self.  (For a field-basis
        orbit this is the reference's result; for other bases the reference relabels the rolled field with the old basis.)"""
        f = self.transform(to="field")
        return f._new(state=np.roll(f.state, shift, axis=axis)).transform(to=self.basis)

    def cell_shift(self, n_cell, axis=0):
        """Rotation by 1 / n_cell of the domain along ``axis``, to the nearest grid point (ks/orbits.py:1196-1232)."""
        shift = np.sign(n_cell) * np.array(self.discretization)[np.array(axis)] // np.abs(n_cell)
        return self.roll(shift, axis=axis)

    def reflection(self, axis=1, signed=True):
        """-u(L - x, t) (ks/orbits.py:1143-1170): flip the columns, roll by one along ``axis`` (the reflection axis sits on a
        grid point), negate; the spatial shift changes sign."""
        f = self.transform(to="field")
        t, x, sft = self.parameters
        return f._new(state=-1.0 * np.roll(np.fliplr(f.state), 1, axis=axis), parameters=(t, x, -1 * sft)).transform(to=self.basis)

    def __getitem__(self, key):
        """Slice of the state with the dimensions scaled along (core.py:553-668): an axis cut to a fraction of its points
        keeps that fraction of its period / length, an axis cut to one point gets dimension 0; only basic indexing that keeps
        both axes is allowed."""
        try:
            cut = self.state[key]
        except IndexError as ie:
            raise ValueError("attempting to slice an Orbit whose state is empty." if self.size == 0 else "invalid slice") from ie
        if cut.ndim != self.state.ndim:
            raise ValueError(f"{type(self).__name__}.__getitem__ does not support arguments which squeeze, flatten, or otherwise "
                             "reduce the dimension of the state array")
        if self.parameters is None:
            return self._new(state=cut, discretization=None)
        dims = tuple(float(d) * (new / old) if new > 1 else 0.0 for d, new, old in zip(self.parameters[:2], cut.shape, self.shape))
        return self._new(state=cut, parameters=dims + tuple(self.parameters[2:]), discretization=None)

    def to_fundamental_domain(self, **kwargs):
        return self  # core.py: no discrete symmetry, the whole tile

    def from_fundamental_domain(self, **kwargs):
        return self

    def concat(self, other, axis=0):
        """Join two orbits (core.py:1282-1380): states concatenated along ``axis`` as they are (the bases are not checked),
        the dimension along ``axis`` summed and the other averaged (glue_dimensions, core.py:899-945), the remaining
        parameters those of self, discretization re-read from the new state."""
        state = np.concatenate((self.state, other.state), axis=axis)
        if self.parameters is not None and other.parameters is not None:
            dims = tuple((2 if i == axis else 1) * 0.5 * (self.parameters[i] + other.parameters[i]) for i in range(2))
        elif self.parameters is not None or other.parameters is not None:
            src = self if self.parameters is not None else other
            dims = tuple(float(d) * (new / old) if new > 1 else 0.0 for d, new, old in zip(src.parameters[:2], state.shape, src.shape))
        else:
            return self._new(state=state, discretization=None)
        base = self.parameters if self.parameters is not None else other.parameters
        return self._new(state=state, parameters=dims + tuple(base[2:]), discretization=None)

    def group_orbit(self, **kwargs):
        """Generator over the group orbit (ks/orbits.py:1355-1401): ``discrete=True`` -- {1, reflection, half-cell shift in
        space, both} x {1, half-cell shift in time}; ``continuous=True`` -- all rolls by multiples of ``rolls``; default --
        those rolls of the orbit and of its reflection."""
        fd = kwargs.get("fundamental_domain", False)
        if kwargs.get("discrete", False):
            half = self.cell_shift(2, axis=1)
            for g_space in (self, self.reflection(), half, half.reflection()):
                for g in (g_space, g_space.cell_shift(2, axis=0)):
                    yield g.to_fundamental_domain() if fd else g
            return
        rolls = kwargs.get("rolls", (1, 1))
        for g in ((self,) if kwargs.get("continuous", False) else (self, self.reflection())):
            for N in range(0, g.n, rolls[0]):
                for M in range(0, g.m, rolls[1]):
                    r = g.roll(N, axis=0).roll(M, axis=1)
                    yield r.to_fundamental_domain() if fd else r


class RelativeOrbitKS(OrbitKS):
    _cls_name = "RelativeOrbitKS"

    def __init__(self, state=None, basis=None, parameters=None, discretization=None, constraints=None, frame="comoving",
                 **kwargs):
        self.frame = frame
        super().__init__(state=state, basis=basis, parameters=parameters, discretization=discretization,
                         constraints=constraints, **kwargs)

    @classmethod
    def _default_parameter_ranges(cls):
        return {"t": (20, 200), "x": (20, 100), "s": (0, 0)}

    @classmethod
    def _default_constraints(cls):
        return {"t": False, "x": False, "s": False}

    def periodic_dimensions(self):
        return (True, True) if self.frame == "comoving" else (False, True)

    def _pad(self, size, axis=0, **kwargs):
        assert self.frame == "comoving", "Mode padding requires comoving frame; set padding=False if plotting"
        return super()._pad(size, axis=axis)

    def _truncate(self, size, axis=0):
        assert self.frame == "comoving", "Mode truncation requires comoving frame; set padding=False if plotting"
        return super()._truncate(size, axis=axis)

    def change_reference_frame(self, frame):
        """Co-moving <-> physical frame (ks/orbits.py:2770-2833): every time row of the spatial modes is rotated by the phase
        q_k * (+-S / T) * t, t running from T (first row) down to 0 (last row); S is always the shift from the co-moving to the
        physical frame."""
        if frame not in ("comoving", "physical"):
            raise ValueError("Trying to change to unrecognizable reference frame.")
        if frame == self.frame:
            return self
        shift = self.s if frame == "physical" else -1.0 * self.s
        sm = self.transform(to="spatial_modes").state
        k = int(self.m // 2) - 1
        t = np.flipud(np.linspace(0, self.t, num=self.n, endpoint=True)).reshape(-1, 1)
        q = (2 * np.pi / self.x) * np.arange(1, k + 1).reshape(1, -1)
        theta = (shift / self.t) * t * q
        c, sn = np.cos(theta), np.sin(theta)
        re, im = sm[:, :k], sm[:, k:]
        rotated = np.concatenate((c * re + sn * im, -sn * re + c * im), axis=1)
        return self._new(state=rotated, basis="spatial_modes", frame=frame).transform(to=self.basis)

    def to_fundamental_domain(self, **kwargs):
        return self.change_reference_frame("physical")  # ks/orbits.py:3013-3014

    def from_fundamental_domain(self, **kwargs):
        return self.change_reference_frame("comoving")  # ks/orbits.py:3010-3011

    def populate(self, attr="all", **kwargs):
        # the reference resets S to 0 unless it estimates a shift from an existing state (ks/orbits.py:2872-2883);
        # shift estimation (calculate_spatial_shift, fsolve on the host) is outside the hot path.
        super().populate(attr=attr, **kwargs)
        self.parameters = (self.parameters[0], self.parameters[1], 0.0)
        return self


class _HalfColumnMixin:
    @staticmethod
    def _pad_space(state, m, size):  # one block of k columns: zeros appended (ks/orbits.py:3346-3350, :3756-3760)
        return np.pad(state, ((0, 0), (0, (size - m) // 2)))

    @staticmethod
    def _truncate_space(state, m, keep):  # (ks/orbits.py:3391-3394, :3801-3804)
        return state[:, :keep]

    def _mode_discretization(self, shape):
        return shape[0] + 1, 2 * shape[1] + 2


class ShiftReflectionOrbitKS(_HalfColumnMixin, OrbitKS):
    _cls_name = "ShiftReflectionOrbitKS"

    def to_fundamental_domain(self, half=0, **kwargs):
        """One half of the period (ks/orbits.py:3688-3700): the last n/2 rows for half 0, the first n/2 otherwise."""
        f, h = self.transform(to="field"), int(self.n // 2)
        return (f[-h:, :] if half == 0 else f[:-h, :]).transform(to=self.basis)

    def from_fundamental_domain(self, half=0, **kwargs):
        """The full period from one half and its reflection (ks/orbits.py:3702-3711)."""
        f = self.transform(to="field")
        joined = f.reflection().concat(f, axis=0) if half == 0 else f.concat(f.reflection(), axis=0)
        return joined.transform(to=self.basis)

    def concat(self, other, axis=0):
        """ks/orbits.py:3640-3676: joined in space, the result is rolled by half of the second tile's width."""
        joined = super().concat(other, axis=axis)
        return joined.roll(other.m // 2, axis=1) if axis == 1 else joined


class AntisymmetricOrbitKS(_HalfColumnMixin, OrbitKS):
    _cls_name = "AntisymmetricOrbitKS"

    def to_fundamental_domain(self, half=0, **kwargs):
        """One half of the domain (ks/orbits.py:3240-3252): the first m/2 columns for half 0, the last m/2 otherwise."""
        f, h = self.transform(to="field"), int(self.m // 2)
        return (f[:, :-h] if half == 0 else f[:, -h:]).transform(to=self.basis)

    def from_fundamental_domain(self, half=0, **kwargs):
        """The full domain from one half and its reflection (ks/orbits.py:3229-3238)."""
        f = self.transform(to="field")
        joined = f.concat(f.reflection(), axis=1) if half == 0 else f.reflection().concat(f, axis=1)
        return joined.transform(to=self.basis)

    @staticmethod
    def _default_shape():
        return 1, 32

    @classmethod
    def _default_parameter_ranges(cls):
        return {"t": (20.0, 200.0), "x": (38.5, 100.0)}


CLASSES = {c._cls_name: c for c in (OrbitKS, RelativeOrbitKS, ShiftReflectionOrbitKS, AntisymmetricOrbitKS)}
