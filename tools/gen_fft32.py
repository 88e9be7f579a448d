"""Generator of the straight-line length-32 real FFTs of the 32x32 kernel:  python tools/gen_fft32.py
writes orbithunter_b200/csrc/ks_fft32_gen.cuh.

Every value the generator tracks is `scale * register` with a compile-time `scale`.  A linear combination
ca * a + cb * b of two tracked values is ONE instruction, fma(r, b_reg, a_reg) with r = cb * b.scale / (ca * a.scale) (an
add or subtract when r = +-1), whose result carries the pending scale ca * a.scale.  Twiddle multiplications therefore never
cost an instruction of their own: they end up in the constants of later fmas.  The operand that keeps the pending scale at
+-1 is preferred as the unit one; with a DC term in the spectrum every output then ends with pending scale exactly +-1 (the
DC chain has unit coefficients all the way down), so the outputs are true values and no multiply is ever issued.
  * inverse (Hermitian half spectrum -> 32 real samples): decimation in frequency, the transpose of the forward algorithm:
    E[k] = Y[k] + conj Y[N/2-k] feeds the even samples, O[k] = (Y[k] - conj Y[N/2-k]) W^-k the odd ones.  6 instructions per
    interior k and level: 162 in all, against 233 for the route through a packed complex length-16 transform.
  * forward (32 real samples -> half spectrum): decimation in time, 163 (the hand-written ksfft::rfft_dit: 164).
The spectra follow the conventions of ks_warp32x1.cuh: unnormalised, forward sign e^{-i...}, Nyquist term dropped.
Reference call sites replaced: scipy.fft.rfft / irfft at orbithunter/ks/orbits.py:2245, :2307, :2352, :2409.
"""
import math
import os

import numpy as np

N = 32


class V:
    __slots__ = ("ref", "scale")

    def __init__(self, ref, scale=1.0):
        self.ref, self.scale = ref, scale


class Gen:
    def __init__(self):
        self.ops = []  # (dst id, kind, const, b ref, a ref):  dst = const * b + a | a + b | a - b | const * a
        self.count = 0

    def lin(self, a, ca, b, cb):
        if a is None and b is None:
            return None
        if a is None:
            return V(b.ref, cb * b.scale)
        if b is None:
            return V(a.ref, ca * a.scale)
        A, B = ca * a.scale, cb * b.scale
        if abs(A) != 1.0 and abs(B) == 1.0:  # keep the pending scale at +-1 whenever one operand allows it
            a, b, A, B = b, a, B, A
        r = B / A
        dst = ("t", self.count)
        self.count += 1
        kind = "add" if r == 1.0 else "sub" if r == -1.0 else "fma"
        self.ops.append((dst, kind, r, b.ref, a.ref))
        return V(dst, A)

    def resolve(self, v):
        """one multiply: the true value in a register (scale 1)"""
        if v is None or abs(v.scale) == 1.0:  # a sign is folded into the consumer (operand negation is free)
            return v
        dst = ("t", self.count)
        self.count += 1
        self.ops.append((dst, "mul", v.scale, None, v.ref))
        return V(dst, 1.0)

    # ---- the two transforms -------------------------------------------------------------------------------------------
    def c2r(self, Yr, Yi, n):
        """Yr, Yi: lists of length n/2 + 1 (None = structurally zero) -> list of n values x[j] = sum_k Y[k] e^{+2 pi i k j / n}"""
        if n == 1:
            return [Yr[0]]
        if n == 2:
            return [self.lin(Yr[0], 1.0, Yr[1], 1.0), self.lin(Yr[0], 1.0, Yr[1], -1.0)]
        half, q = n // 2, n // 4
        Er, Ei, Or, Oi = [None] * (q + 1), [None] * (q + 1), [None] * (q + 1), [None] * (q + 1)
        Er[0] = self.lin(Yr[0], 1.0, Yr[half], 1.0)
        Or[0] = self.lin(Yr[0], 1.0, Yr[half], -1.0)
        for k in range(1, q):
            Er[k] = self.lin(Yr[k], 1.0, Yr[half - k], 1.0)
            Ei[k] = self.lin(Yi[k], 1.0, Yi[half - k], -1.0)
            Dr = self.lin(Yr[k], 1.0, Yr[half - k], -1.0)
            Di = self.lin(Yi[k], 1.0, Yi[half - k], 1.0)
            c, s = math.cos(2 * math.pi * k / n), math.sin(2 * math.pi * k / n)
            Or[k] = self.lin(Dr, c, Di, -s)
            Oi[k] = self.lin(Di, c, Dr, s)
        Er[q] = None if Yr[q] is None else V(Yr[q].ref, 2.0 * Yr[q].scale)
        Or[q] = None if Yi[q] is None else V(Yi[q].ref, -2.0 * Yi[q].scale)
        xe, xo = self.c2r(Er, Ei, half), self.c2r(Or, Oi, half)
        out = [None] * n
        out[0::2], out[1::2] = xe, xo
        return out

    def r2c(self, x):
        """x: list of n values -> (Yr, Yi) of length n/2 + 1, Y[k] = sum_j x[j] e^{-2 pi i k j / n}"""
        n = len(x)
        if n == 1:
            return [x[0]], [None]
        if n == 2:
            return [self.lin(x[0], 1.0, x[1], 1.0), self.lin(x[0], 1.0, x[1], -1.0)], [None, None]
        half, q = n // 2, n // 4
        Er, Ei = self.r2c(x[0::2])
        Or, Oi = self.r2c(x[1::2])
        Yr, Yi = [None] * (half + 1), [None] * (half + 1)
        Yr[0] = self.lin(Er[0], 1.0, Or[0], 1.0)
        Yr[half] = self.lin(Er[0], 1.0, Or[0], -1.0)
        Yr[q] = Er[q]
        Yi[q] = None if Or[q] is None else V(Or[q].ref, -Or[q].scale)
        for k in range(1, q):
            c, s = math.cos(2 * math.pi * k / n), math.sin(2 * math.pi * k / n)
            Pr = self.lin(Or[k], c, Oi[k], s)
            Pi = self.lin(Oi[k], c, Or[k], -s)
            Yr[k] = self.lin(Er[k], 1.0, Pr, 1.0)
            Yi[k] = self.lin(Ei[k], 1.0, Pi, 1.0)
            Yr[half - k] = self.lin(Er[k], 1.0, Pr, -1.0)
            Yi[half - k] = self.lin(Pi, 1.0, Ei[k], -1.0)
        return Yr, Yi

    # ---- output ---------------------------------------------------------------------------------------------------------
    def live_ops(self, outputs):
        need = {v.ref for v in outputs if v is not None and v.ref[0] == "t"}
        keep = []
        for op in reversed(self.ops):
            dst, kind, r, b, a = op
            if dst in need:
                keep.append(op)
                for ref in (a, b):
                    if ref is not None and ref[0] == "t":
                        need.add(ref)
        return list(reversed(keep))

    def evaluate(self, ops, inputs):
        val = dict(inputs)
        for dst, kind, r, b, a in ops:
            if kind == "add":
                val[dst] = val[a] + val[b]
            elif kind == "sub":
                val[dst] = val[a] - val[b]
            elif kind == "mul":
                val[dst] = r * val[a]
            else:
                val[dst] = r * val[b] + val[a]
        return val


def cref(ref):
    return f"t{ref[1]}" if ref[0] == "t" else ref[1]


def cval(v):
    """C expression of a value whose pending scale is +-1"""
    assert abs(v.scale) == 1.0, v.scale
    return cref(v.ref) if v.scale > 0 else "-" + cref(v.ref)


CONSTANTS = []  # distinct |r| of the fmas, shared by all generated functions (KS_FFT32_CONST_BANK build: c[bank][offset] operands)


def cconst(r):
    mag = abs(r)
    if mag not in CONSTANTS:
        CONSTANTS.append(mag)
    return f"{'-' if r < 0 else ''}KS_FFT32_C({CONSTANTS.index(mag)}, {mag!r})"


def by_level(ops):
    """the same operations ordered by their depth in the dependency graph (independent operations next to each other)"""
    depth = {}
    for dst, kind, r, b, a in ops:
        depth[dst] = 1 + max([depth.get(ref, 0) for ref in (a, b) if ref is not None and ref[0] == "t"] + [0])
    return sorted(ops, key=lambda op: depth[op[0]])  # stable: generation order within a level


def by_height(ops):
    """list schedule: among the operations whose operands are available, the one with the longest path to an output first"""
    users, height = {}, {}
    for op in ops:
        for ref in (op[4], op[3]):
            if ref is not None and ref[0] == "t":
                users.setdefault(ref, []).append(op[0])
    for op in reversed(ops):
        height[op[0]] = 1 + max([height[u] for u in users.get(op[0], [])] + [0])
    done, out, pending = set(), [], list(ops)
    remaining_uses = {k: len(v) for k, v in users.items()}
    frees = os.environ.get("KS_FFT32_TIE_BREAK") == "last_use"  # tie-break on operations that end an operand's life (no gain)
    while pending:
        ready = [op for op in pending if all(ref is None or ref[0] != "t" or ref in done for ref in (op[4], op[3]))]

        def key(op):
            last = sum(1 for ref in (op[4], op[3]) if ref is not None and ref[0] == "t" and remaining_uses.get(ref, 0) == 1)
            return (height[op[0]], last) if frees else (height[op[0]],)

        best = max(ready, key=key)
        out.append(best)
        done.add(best[0])
        pending.remove(best)
        for ref in (best[4], best[3]):
            if ref is not None and ref[0] == "t":
                remaining_uses[ref] -= 1
    return out


def by_latency(ops, lat):
    """list schedule against a latency model: an operation may issue `lat` slots after its operands; among those that may,
    the one with the longest path to an output goes first; if none may, the one that becomes available soonest"""
    users, height = {}, {}
    for op in ops:
        for ref in (op[4], op[3]):
            if ref is not None and ref[0] == "t":
                users.setdefault(ref, []).append(op[0])
    for op in reversed(ops):
        height[op[0]] = 1 + max([height[u] for u in users.get(op[0], [])] + [0])
    issued, out, pending, slot = {}, [], list(ops), 0
    while pending:
        avail = []
        for op in pending:
            deps = [ref for ref in (op[4], op[3]) if ref is not None and ref[0] == "t"]
            if all(d in issued for d in deps):
                avail.append((max([issued[d] + lat for d in deps] + [0]), op))
        ready = [op for t, op in avail if t <= slot]
        if ready:
            best = max(ready, key=lambda op: height[op[0]])
        else:
            best = min(avail, key=lambda to: (to[0], -height[to[1][0]]))[1]
        out.append(best)
        issued[best[0]] = slot
        pending.remove(best)
        slot += 1
    return out


def emit_body(ops):
    lines = []
    # ptxas keeps much of the source order: recursion order 334 us per 65 536 orbits, dependency-level order 327 us, longest
    # path first 320.5 us (profiles/r02z_x1_fft_emission_order.jsonl)
    order = os.environ.get("KS_FFT32_ORDER", "height")
    if order.startswith("latency"):
        ops = by_latency(ops, int(order[7:] or 4))
    elif order == "height":
        ops = by_height(ops)
    elif order == "level":
        ops = by_level(ops)
    for dst, kind, r, b, a in ops:
        if kind == "add":
            e = f"{cref(a)} + {cref(b)}"
        elif kind == "sub":
            e = f"{cref(a)} - {cref(b)}"
        elif kind == "mul":
            e = f"{cconst(r)} * {cref(a)}"
        else:
            e = f"fma({cconst(r)}, {cref(b)}, {cref(a)})"
        lines.append(f"  const double t{dst[1]} = {e};")
    return lines


def table(name, values):
    body = ", ".join(repr(float(v)) for v in values)
    return f"static constexpr double {name}[{len(values)}] = {{{body}}};"


def spectrum_inputs(with_dc):
    Yr = [V(("in", "h0")) if with_dc else None] + [V(("in", f"hr[{k}]")) for k in range(1, 16)] + [None]
    Yi = [None] + [V(("in", f"hi[{k}]")) for k in range(1, 16)] + [None]
    return Yr, Yi


def build_inverse(with_dc, resolved):
    g = Gen()
    x = g.c2r(*spectrum_inputs(with_dc), N)
    if resolved:
        x = [g.resolve(v) for v in x]
    ops = g.live_ops(x)
    return g, ops, x


def build_forward(in_scales, resolved):
    g = Gen()
    x = [V(("in", f"x[{j}]"), float(in_scales[j])) for j in range(N)]
    Yr, Yi = g.r2c(x)
    outs = [Yr[k] for k in range(1, 16)] + [Yi[k] for k in range(1, 16)]
    if resolved:
        outs = [g.resolve(v) for v in outs]
    ops = g.live_ops(outs)
    return g, ops, outs


def check_inverse(g, ops, x, with_dc):
    rng = np.random.RandomState(1)
    h0, hr, hi = rng.randn(), rng.randn(16), rng.randn(16)
    inputs = {("in", "h0"): h0}
    for k in range(1, 16):
        inputs[("in", f"hr[{k}]")] = hr[k]
        inputs[("in", f"hi[{k}]")] = hi[k]
    val = g.evaluate(ops, inputs)
    Y = np.zeros(N, dtype=complex)
    Y[0] = h0 if with_dc else 0.0
    for k in range(1, 16):
        Y[k] = hr[k] + 1j * hi[k]
        Y[N - k] = np.conj(Y[k])
    ref = (np.fft.ifft(Y) * N).real
    got = np.array([v.scale * val[v.ref] for v in x])
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < 1e-14, err
    return err


def check_forward(g, ops, outs, in_scales):
    rng = np.random.RandomState(2)
    xr = rng.randn(N)
    val = g.evaluate(ops, {("in", f"x[{j}]"): xr[j] for j in range(N)})
    ref = np.fft.fft(np.asarray(in_scales) * xr)
    got = np.array([v.scale * val[v.ref] for v in outs])
    want = np.concatenate((ref[1:16].real, ref[1:16].imag))
    err = np.max(np.abs(got - want)) / np.max(np.abs(want))
    assert err < 1e-14, err
    return err


def count(ops):
    return len(ops), sum(1 for o in ops if o[1] == "fma"), sum(1 for o in ops if o[1] in ("add", "sub")), sum(1 for o in ops if o[1] == "mul")


def main():
    out = ["// GENERATED by tools/gen_fft32.py -- do not edit.  Straight-line length-32 real FFTs with deferred (pending) scales;",
           "// see the generator for the method and the conventions.  Instruction counts (fp64, as generated):"]
    funcs, notes = [], []
    # inverse: DC + 15 complex coefficients -> 32 true samples
    g, ops, x = build_inverse(True, True)
    check_inverse(g, ops, x, True)
    notes.append(("irfft32", count(ops)))
    funcs += ["// x[j] = h0 + sum_{k=1..15} 2 Re((hr[k] + i hi[k]) e^{+2 pi i k j / 32})",
              "KS_DEVI void irfft32(double h0, const double (&hr)[16], const double (&hi)[16], double (&x)[32]) {"] + emit_body(ops)
    funcs += [f"  x[{j}] = {cval(v)};" for j, v in enumerate(x)] + ["}", ""]
    # forward: 32 true samples -> DC and k = 1..15
    g = Gen()
    xin = [V(("in", f"x[{j}]")) for j in range(N)]
    Yr, Yi = g.r2c(xin)
    outs = [g.resolve(v) for v in [Yr[0]] + [Yr[k] for k in range(1, 16)] + [Yi[k] for k in range(1, 16)]]
    ops = g.live_ops(outs)
    rng = np.random.RandomState(3)
    xr = rng.randn(N)
    val = g.evaluate(ops, {("in", f"x[{j}]"): xr[j] for j in range(N)})
    ref = np.fft.fft(xr)
    got = np.array([v.scale * val[v.ref] for v in outs])
    want = np.concatenate(([ref[0].real], ref[1:16].real, ref[1:16].imag))
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-14
    notes.append(("rfft32", count(ops)))
    funcs += ["// c0 = sum_j x[j];  cr[k] + i ci[k] = sum_j x[j] e^{-2 pi i k j / 32}, k = 1..15 (cr[0] = ci[0] = 0)",
              "KS_DEVI void rfft32(const double (&x)[32], double& c0, double (&cr)[16], double (&ci)[16]) {"] + emit_body(ops)
    funcs += [f"  c0 = {cval(outs[0])};", "  cr[0] = 0.0;", "  ci[0] = 0.0;"]
    funcs += [f"  cr[{k + 1}] = {cval(v)};" for k, v in enumerate(outs[1:16])]
    funcs += [f"  ci[{k + 1}] = {cval(v)};" for k, v in enumerate(outs[16:])] + ["}", ""]
    for name, (tot, nf, na, nm) in notes:
        out.append(f"//   {name:10s} {tot:4d}  ({nf} fma, {na} add/sub, {nm} mul)")
    consts = ["// fma constants: read from a __constant__ table (uniform loads LDCU of the constant bank) instead of literals, which",
              "// ptxas materialises as UMOV pairs into uniform registers: 560 fewer instructions issued per evaluation of the 32x32",
              "// kernel, same fp64 work (-DKS_FFT32_LITERALS restores the literals; measured equal within noise, profiles/r02s_*)",
              "#if !defined(KS_FFT32_LITERALS) && !defined(KS_HOST_EMUL)",
              f"static __constant__ double kFft32C[{len(CONSTANTS)}] = {{{', '.join(repr(c) for c in CONSTANTS)}}};",
              "#define KS_FFT32_C(i, v) kFft32C[i]", "#else", "#define KS_FFT32_C(i, v) (v)", "#endif", ""]
    out += ["#pragma once", '#include "ks_platform.cuh"', "", "namespace ksg32 {", ""] + consts + funcs + ["#undef KS_FFT32_C", "}  // namespace ksg32", ""]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orbithunter_b200", "csrc", "ks_fft32_gen.cuh")
    with open(path, "w") as fh:
        fh.write("\n".join(out))
    for name, c in notes:
        print(name, c)


if __name__ == "__main__":
    main()
