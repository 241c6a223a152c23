"""TEST INFRASTRUCTURE ONLY -- run the reference's own renderer source without TensorFlow.

The reference (TF 1.12 graph code) cannot be imported here: TensorFlow and
tensorpack are absent and un-installable.  But the hot path is ~120 lines of
pure tensor algebra that use ~25 ``tf.*`` names.  This module

1. provides ``tf``: a namespace with exactly those names, each mapped onto the
   torch op with the same semantics (shapes, axis conventions, dtypes), and
2. ``load_reference_functions(path, names, K)``: parses a reference file with
   ``ast``, compiles ONLY the named top-level ``def``s (nothing else in the file
   is executed -- the module-level tensorpack imports never run) and binds them
   to a namespace holding the shim, ``np`` and the module global ``nb_landmarks``
   the functions read (mask/mask_gen_dynamic.py:126-128, 259-264).

The reference source is read from ``/root/reference`` at call time and is never
copied into this repository.  Used by ``tests/golden/make_golden.py`` to create
the committed golden vectors and by ``tests/test_oracle.py`` (skipped
when ``/root/reference`` is absent, e.g. on the GPU box).

What this pins and what it does not: op ORDER, transposes, gather indices,
broadcast shapes, sign flips and cast boundaries are the reference's own.  The
primitive kernels (inv, det, matmul, exp) are torch's, not TF/Eigen's; float32
``cos``/``sin`` and the float32 3x3 ``matmul`` are pinned by definition (correctly rounded trig,
fixed summation order) -- this module carries its OWN numpy implementations of them, of ``linspace`` and
of ``l2_normalize``; tests/test_oracle.py asserts they equal ``reference_port``'s bit for bit.
"""
from __future__ import annotations

import ast
import contextlib
import types
from typing import Dict, Iterable

import numpy as np
import torch


# ---- the shim's OWN definitions of the three primitives whose last-ulp behaviour matters (numpy, written
# independently of reference_port's torch versions; tests/test_oracle.py asserts the two agree bit for bit and checks
# the trig against mpmath's correctly rounded values).  Not imported from the port: a checker that borrows the
# checked code's primitives pins nothing.
class _TrigF32(torch.autograd.Function):
    """float32 cos / sin = the float64 libm value of the float32 argument, rounded to float32 (numpy forward; the
    derivative is evaluated the same way)."""

    @staticmethod
    def forward(ctx, a, which):
        x = a.detach().cpu().numpy().astype(np.float64)
        ctx.save_for_backward(a)
        ctx.which = which
        val = np.cos(x) if which == "cos" else np.sin(x)
        return torch.from_numpy(val.astype(np.float32)).to(a.device)

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        x = a.detach().cpu().numpy().astype(np.float64)
        d = -np.sin(x) if ctx.which == "cos" else np.cos(x)
        return (g.to(torch.float64) * torch.from_numpy(d).to(g.device)).to(torch.float32), None


def cos_f32(a: torch.Tensor) -> torch.Tensor:
    return _TrigF32.apply(a, "cos") if a.dtype == torch.float32 else torch.cos(a)


def sin_f32(a: torch.Tensor) -> torch.Tensor:
    return _TrigF32.apply(a, "sin") if a.dtype == torch.float32 else torch.sin(a)


def _mm3_np(A: np.ndarray, Bm: np.ndarray, bt: bool = False, at: bool = False) -> np.ndarray:
    """3x3 float32 product with explicit loops: every multiply and add rounded to float32, order ((k=0 + k=1) + k=2)."""
    shape = np.broadcast_shapes(A.shape, Bm.shape)
    A = np.broadcast_to(A, shape).astype(np.float32)
    Bm = np.broadcast_to(Bm, shape).astype(np.float32)
    out = np.empty(shape, np.float32)
    a_ = (lambda i, k: A[..., k, i]) if at else (lambda i, k: A[..., i, k])
    b_ = (lambda k, j: Bm[..., j, k]) if bt else (lambda k, j: Bm[..., k, j])
    for i in range(3):
        for j in range(3):
            p0 = (a_(i, 0) * b_(0, j)).astype(np.float32)
            p1 = (a_(i, 1) * b_(1, j)).astype(np.float32)
            p2 = (a_(i, 2) * b_(2, j)).astype(np.float32)
            out[..., i, j] = ((p0 + p1).astype(np.float32) + p2).astype(np.float32)
    return out


class _MatMul3F32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return torch.from_numpy(_mm3_np(a.detach().cpu().numpy(), b.detach().cpu().numpy())).to(a.device)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        gn, an, bn = (t.detach().cpu().numpy() for t in (g, a, b))
        ga = torch.from_numpy(_mm3_np(gn, bn, bt=True))          # g b^T
        gb = torch.from_numpy(_mm3_np(an, gn, at=True))          # a^T g
        while ga.dim() > a.dim():
            ga = ga.sum(0)
        while gb.dim() > b.dim():
            gb = gb.sum(0)
        for d in range(a.dim()):
            if a.shape[d] == 1 and ga.shape[d] != 1:
                ga = ga.sum(d, keepdim=True)
            if b.shape[d] == 1 and gb.shape[d] != 1:
                gb = gb.sum(d, keepdim=True)
        return ga.to(a.device), gb.to(b.device)


def matmul3_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _MatMul3F32.apply(a, b)


def tf_linspace_f32(n: int) -> np.ndarray:
    """TF 1.12 LinSpace in float32: out[i] = start + step * i, step = (stop - start) / (n - 1), scalar float32 ops."""
    start, stop = np.float32(-1.0), np.float32(1.0)
    if n == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32(np.float32(stop - start) / np.float32(n - 1))
    return np.array([np.float32(start + np.float32(step * np.float32(i))) for i in range(n)], dtype=np.float32)


def _dt(d):
    return d


def _stack(values, axis=0, name=None):
    return torch.stack([torch.as_tensor(v) for v in values], dim=axis)


def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    if a.dtype == torch.float32 and a.shape[-2:] == (3, 3) and b.shape[-2:] == (3, 3):
        return matmul3_f32(a, b)      # pinned evaluation order, see reference_port.matmul3_f32
    return torch.matmul(a, b)


def _constant(value, dtype=None, name=None):
    return torch.tensor(np.asarray(value), dtype=dtype)


def _gather(params, indices, axis=0, name=None):
    idx = torch.as_tensor(indices, dtype=torch.long)
    return torch.index_select(params, axis, idx)


def _linspace(start, stop, num, name=None):
    assert float(start) == -1.0 and float(stop) == 1.0, "shim only models the call the path makes"
    return torch.from_numpy(tf_linspace_f32(int(num)))


def _meshgrid(x, y, indexing="xy"):
    assert indexing == "xy"
    X = x.view(1, -1).expand(y.numel(), x.numel())
    Y = y.view(-1, 1).expand(y.numel(), x.numel())
    return [X, Y]


def _reshape(t, shape, name=None):
    return t.reshape([int(s) for s in shape])


def _reduce_sum(t, axis=None, keepdims=False, name=None):
    return t.sum(dim=axis, keepdim=keepdims) if axis is not None else t.sum()


def _reduce_max(t, axis=None, keepdims=False, name=None):
    if isinstance(t, (list, tuple)):
        t = torch.stack(list(t), dim=0)
    return t.max(dim=axis, keepdim=keepdims).values if axis is not None else t.max()


def _zeros(shape, dtype=torch.float32, name=None):
    return torch.zeros([int(s) for s in shape], dtype=dtype)


def _squeeze(t, axis=None, name=None):
    return t.squeeze(axis) if axis is not None else t.squeeze()


def _transpose(t, perm=None, name=None):
    return t.permute(*perm) if perm is not None else t.t()


def _tile(t, multiples, name=None):
    return t.repeat(*[int(m) for m in multiples])


@contextlib.contextmanager
def _name_scope(*a, **k):
    yield


def make_tf() -> types.SimpleNamespace:
    """Build the ``tf`` stand-in.  Only names used on the path exist; anything else raises."""
    linalg = types.SimpleNamespace(
        inv=lambda a, name=None: torch.linalg.inv(a),
        det=lambda a, name=None: torch.linalg.det(a),
        matmul=_matmul,
    )
    tf = types.SimpleNamespace(
        float32=torch.float32, float64=torch.float64, newaxis=None,
        zeros_like=lambda t, name=None: torch.zeros_like(torch.as_tensor(t)),
        ones_like=lambda t, name=None: torch.ones_like(torch.as_tensor(t)),
        zeros=_zeros, stack=_stack, cos=cos_f32, sin=sin_f32, exp=torch.exp,
        matmul=_matmul, constant=_constant, reshape=_reshape,
        shape=lambda t: tuple(t.shape),
        cast=lambda t, dtype, name=None: t.to(dtype),
        to_float=lambda t, name=None: t.to(torch.float32),
        identity=lambda t, name=None: t,
        transpose=_transpose, gather=_gather, squeeze=_squeeze, tile=_tile,
        expand_dims=lambda t, axis, name=None: t.unsqueeze(axis),
        linspace=_linspace, meshgrid=_meshgrid, reduce_sum=_reduce_sum, reduce_max=_reduce_max,
        name_scope=_name_scope, linalg=linalg,
    )
    return tf


def load_reference_functions(path: str, names: Iterable[str], nb_landmarks: int,
                             extra_globals: Dict | None = None) -> Dict[str, object]:
    """Compile the named top-level functions of reference file ``path`` against the shim."""
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = set(names)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    # a file may define a name more than once (never the case for the path, but be strict)
    seen = [n.name for n in picked]
    missing = wanted - set(seen)
    if missing:
        raise KeyError(f"{path}: no top-level def for {sorted(missing)}")
    if len(seen) != len(set(seen)):
        raise ValueError(f"{path}: duplicate definitions among {seen}")
    mod = ast.Module(body=picked, type_ignores=[])
    ns: Dict[str, object] = {"tf": make_tf(), "np": np, "nb_landmarks": int(nb_landmarks)}
    if extra_globals:
        ns.update(extra_globals)
    exec(compile(mod, path, "exec"), ns)  # noqa: S102 -- executing the read-only reference, on purpose
    return {n: ns[n] for n in wanted}


# ------------------------------------------------------------------------------------------------
# Geometry tails: Model.get_musigma (:369-403) and Model.transform_mu_sigma (:405-460) are METHODS that mix the
# dense heads (out of scope) with the tensor algebra of rows a6 / a7.  They are executed as they stand, with
# ``tf.layers.dense`` replaced by a queue that hands out pre-drawn head pre-activations (hidden layers get zeros),
# so that everything downstream of the heads -- tanh/sigmoid scalings, l2_normalize, cross products, concat order,
# svd, diag, the Euler stacks, casts and matmuls -- is the reference's own code.
class DenseQueue:
    """Stand-in for ``tf.layers.dense``: returns ``activation(next queued pre-activation)``; layers whose width is
    not a head width (the hidden NF*4 layers) return zeros."""

    def __init__(self, head_widths, preacts):
        self.head_widths = set(int(w) for w in head_widths)
        self.preacts = list(preacts)
        self.calls = []

    def __call__(self, x, units, activation=None, name=None):
        rows = x.shape[0]
        units = int(units)
        if units in self.head_widths:
            p = self.preacts.pop(0)
            assert tuple(p.shape) == (rows, units), (tuple(p.shape), rows, units)
            self.calls.append(units)
            return activation(p) if activation is not None else p
        return torch.zeros(rows, units, dtype=torch.float32)


def _svd(a, full_matrices=False, compute_uv=True, name=None):
    U, S, Vh = torch.linalg.svd(a, full_matrices=full_matrices)
    return S, U, Vh.transpose(-1, -2)          # TensorFlow's order: s, u, v


def _l2_normalize(x, axis=-1, epsilon=1e-12, name=None):
    """``tf.math.l2_normalize`` (TF 1.12 nn_impl.py): x * rsqrt(maximum(reduce_sum(square(x), axis, keepdims), epsilon))."""
    sq = torch.sum(torch.square(x), dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.maximum(sq, torch.tensor(epsilon, dtype=x.dtype)))


def make_tf_geometry(dense: DenseQueue) -> types.SimpleNamespace:
    tf = make_tf()
    tf.layers = types.SimpleNamespace(dense=dense)
    tf.nn = types.SimpleNamespace(tanh=torch.tanh, sigmoid=torch.sigmoid,
                                  leaky_relu=lambda t: torch.nn.functional.leaky_relu(t, 0.2))
    tf.math = types.SimpleNamespace(l2_normalize=_l2_normalize)
    tf.linalg.cross = lambda a, b, name=None: torch.linalg.cross(a, b, dim=-1)
    tf.linalg.diag = lambda d, name=None: torch.diag_embed(d)
    tf.linalg.svd = _svd
    tf.concat = lambda values, axis, name=None: torch.cat(list(values), dim=axis)
    tf.sqrt = torch.sqrt
    tf.squeeze = lambda t, axis=None, name=None: t.squeeze(tuple(axis) if isinstance(axis, (list, tuple)) else axis)
    tf.get_variable = lambda name, shape=None, dtype=None, trainable=True, initializer=None: torch.ones(
        [int(v) for v in shape], dtype=dtype)
    tf.ones_initializer = None
    tf.variable_scope = _name_scope
    return tf


def load_reference_methods(path: str, class_name: str, names: Iterable[str], nb_landmarks: int, dense: DenseQueue,
                           NF: int = 64) -> Dict[str, object]:
    """Compile the named methods of class ``class_name`` in reference file ``path`` as plain functions (decorators
    dropped, ``self`` kept as the first parameter) against the geometry shim."""
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name]
    if len(cls) != 1:
        raise KeyError(f"{path}: class {class_name} not found exactly once")
    wanted = set(names)
    picked = [n for n in cls[0].body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    if {n.name for n in picked} != wanted:
        raise KeyError(f"{path}: missing methods {sorted(wanted - {n.name for n in picked})}")
    for n in picked:
        n.decorator_list = []
    mod = ast.Module(body=picked, type_ignores=[])
    ns: Dict[str, object] = {"tf": make_tf_geometry(dense), "np": np, "nb_landmarks": int(nb_landmarks), "NF": int(NF),
                             "argscope": _name_scope, "Conv2D": None, "Conv2DTranspose": None, "INLReLU": None}
    exec(compile(mod, path, "exec"), ns)  # noqa: S102 -- executing the read-only reference, on purpose
    return {n: ns[n] for n in wanted}
