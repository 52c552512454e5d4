"""AutoGP's covariance-kernel DSL and its flattened wire format.

Mirrors the node set of ``AutoGP.GP`` that the reference re-exports through ``GPConfig``
(`/root/reference/src/NowcastAutoGP.jl:9`; integer codes from
`/root/reference/docs/src/vignettes/setting-priors.md:229-236`). A particle's kernel travels to the
device as a post-order byte program plus a flat ``theta`` vector — docs/KERNEL_SPEC.md §1.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

OP_CONSTANT, OP_LINEAR, OP_SQEXP, OP_GAMMAEXP, OP_PERIODIC, OP_PLUS, OP_TIMES, OP_CHANGEPOINT = range(1, 9)
N_PARAM = {OP_CONSTANT: 1, OP_LINEAR: 3, OP_SQEXP: 2, OP_GAMMAEXP: 3, OP_PERIODIC: 3,
           OP_PLUS: 0, OP_TIMES: 0, OP_CHANGEPOINT: 2}
MAX_PROG = 64
MAX_STACK = 16


class Node:
    """Base class of kernel-tree nodes."""

    def __add__(self, other: "Node") -> "Plus":
        return Plus(self, other)

    def __mul__(self, other: "Node") -> "Times":
        return Times(self, other)


@dataclass(frozen=True)
class Constant(Node):
    value: float


@dataclass(frozen=True)
class Linear(Node):
    intercept: float
    bias: float = 1.0
    amplitude: float = 1.0


@dataclass(frozen=True)
class SquaredExponential(Node):
    lengthscale: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class GammaExponential(Node):
    lengthscale: float
    gamma: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class Periodic(Node):
    lengthscale: float
    period: float
    amplitude: float = 1.0


@dataclass(frozen=True)
class Plus(Node):
    left: Node
    right: Node


@dataclass(frozen=True)
class Times(Node):
    left: Node
    right: Node


@dataclass(frozen=True)
class ChangePoint(Node):
    left: Node
    right: Node
    location: float
    scale: float


_LEAF = {Constant: (OP_CONSTANT, ("value",)),
         Linear: (OP_LINEAR, ("intercept", "bias", "amplitude")),
         SquaredExponential: (OP_SQEXP, ("lengthscale", "amplitude")),
         GammaExponential: (OP_GAMMAEXP, ("lengthscale", "gamma", "amplitude")),
         Periodic: (OP_PERIODIC, ("lengthscale", "period", "amplitude"))}
_BY_OP = {op: (cls, fields) for cls, (op, fields) in _LEAF.items()}


def flatten(node: Node) -> Tuple[bytes, List[float]]:
    """Post-order byte program + theta for one kernel tree."""
    prog: List[int] = []
    theta: List[float] = []

    def rec(nd: Node) -> None:
        if type(nd) in _LEAF:
            op, fields = _LEAF[type(nd)]
            prog.append(op)
            theta.extend(float(getattr(nd, f)) for f in fields)
        elif isinstance(nd, (Plus, Times)):
            rec(nd.left)
            rec(nd.right)
            prog.append(OP_PLUS if isinstance(nd, Plus) else OP_TIMES)
        elif isinstance(nd, ChangePoint):
            rec(nd.left)
            rec(nd.right)
            prog.append(OP_CHANGEPOINT)
            theta.extend((float(nd.location), float(nd.scale)))
        else:
            raise TypeError(f"not a kernel node: {nd!r}")

    rec(node)
    check_program(bytes(prog), len(theta))
    return bytes(prog), theta


def unflatten(prog: bytes, theta: Sequence[float]) -> Node:
    """Inverse of :func:`flatten`."""
    stack: List[Node] = []
    pos = 0
    for op in prog:
        if op in _BY_OP:
            cls, fields = _BY_OP[op]
            stack.append(cls(*[float(x) for x in theta[pos:pos + len(fields)]]))
            pos += len(fields)
        elif op in (OP_PLUS, OP_TIMES):
            r, l = stack.pop(), stack.pop()
            stack.append(Plus(l, r) if op == OP_PLUS else Times(l, r))
        elif op == OP_CHANGEPOINT:
            r, l = stack.pop(), stack.pop()
            stack.append(ChangePoint(l, r, float(theta[pos]), float(theta[pos + 1])))
            pos += 2
        else:
            raise ValueError(f"bad opcode {op}")
    if len(stack) != 1 or pos != len(theta):
        raise ValueError("malformed kernel program")
    return stack[0]


def check_program(prog: bytes, n_theta: int) -> int:
    """Validate a program against docs/KERNEL_SPEC.md §1; returns the max stack depth."""
    if not 0 < len(prog) <= MAX_PROG:
        raise ValueError(f"kernel program length {len(prog)} outside 1..{MAX_PROG}")
    sp = need = deepest = 0
    for op in prog:
        if op not in N_PARAM:
            raise ValueError(f"bad opcode {op}")
        need += N_PARAM[op]
        if op <= OP_PERIODIC:
            sp += 1
            deepest = max(deepest, sp)
        else:
            if sp < 2:
                raise ValueError("malformed kernel program (stack underflow)")
            sp -= 1
    if sp != 1 or need != n_theta:
        raise ValueError("malformed kernel program")
    if deepest > MAX_STACK:
        raise ValueError(f"kernel program needs stack depth {deepest} > {MAX_STACK}")
    return deepest


_SLOT_NAMES = {OP_CONSTANT: ("value",), OP_LINEAR: ("intercept", "bias", "amplitude"),
               OP_SQEXP: ("lengthscale", "amplitude"), OP_GAMMAEXP: ("lengthscale", "gamma", "amplitude"),
               OP_PERIODIC: ("lengthscale", "period", "amplitude"), OP_PLUS: (), OP_TIMES: (),
               OP_CHANGEPOINT: ("location", "scale")}


def theta_slot_names(prog: bytes) -> List[str]:
    """Name of every theta slot a program consumes, in order."""
    return [name for op in prog for name in _SLOT_NAMES[op]]


@dataclass
class FlatEnsemble:
    """B flattened kernels packed for one C-ABI call (CSR-style offsets, int64)."""
    prog: np.ndarray       # uint8 [sum len]
    prog_off: np.ndarray   # int64 [B+1]
    theta: np.ndarray      # float64 [sum ntheta]
    theta_off: np.ndarray  # int64 [B+1]
    noise: np.ndarray      # float64 [B]

    @property
    def size(self) -> int:
        return len(self.noise)


def pack_ensemble(kernels: Sequence[Node], noise: Sequence[float]) -> FlatEnsemble:
    progs, thetas = zip(*(flatten(k) for k in kernels))
    prog_off = np.zeros(len(progs) + 1, np.int64)
    theta_off = np.zeros(len(progs) + 1, np.int64)
    np.cumsum([len(p) for p in progs], out=prog_off[1:])
    np.cumsum([len(t) for t in thetas], out=theta_off[1:])
    return FlatEnsemble(
        prog=np.frombuffer(b"".join(progs), np.uint8).copy(),
        prog_off=prog_off,
        theta=np.asarray([x for t in thetas for x in t], np.float64),
        theta_off=theta_off,
        noise=np.asarray(noise, np.float64).copy(),
    )
