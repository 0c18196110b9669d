"""A small stand-in for the Theano API, built on torch (CPU, float64) -- TEST INFRASTRUCTURE ONLY.

Why: the reference's arithmetic lives in Theano, which cannot be installed here, so the
reference's model files could not be run and the oracle was "parity unpinned".  The reference's model files
(`/root/reference/public/GRU.py`, `GRU_Spatial.py`, `BPR.py`, `PRME.py`, `GeoIE.py`) do compile under Python 3
unmodified; what they need is `import theano`.  This module installs `theano`, `theano.tensor`, ... into
`sys.modules` with exactly the API surface those five files use, so that THE REFERENCE'S OWN GRAPH-BUILDING
CODE runs: its `recurrence` functions, its cost assembly, its `T.grad` / `set_subtensor` / `updates` wiring.
What is ours is the evaluator underneath: lazily evaluated expression nodes, `theano.scan` as a Python loop,
`T.grad` as torch autograd, everything in float64.  Theano's own kernels and float32 rounding are therefore
NOT exercised; the model semantics (which h scores which x, what the L2 term covers, what `Unique` sees, which
rows an update writes, pre-update values everywhere) are the reference's, verbatim.

Used by tests/golden/make_ref_golden.py (run in the build container, where /root/reference exists) to write
tests/golden/ref_*.npz, against which the oracle and the CUDA engine are checked.  Nothing in the product
path imports this file.

Semantics implemented (and the Theano behaviour each follows):
  * theano.shared / get_value / set_value; theano.function(inputs, outputs, updates, givens): all update
    expressions are evaluated from pre-call values, then assigned (Theano's update semantics);
  * theano.scan(fn, sequences, outputs_info, n_steps): fn is called ONCE with symbolic slices (as Theano does)
    and the inner graph replayed n_steps times; sequences longer than n_steps are truncated; outputs_info None
    marks a non-recurrent output;
  * T.grad(cost, wrt) for shared variables and for intermediate nodes (BPR.py:229 differentiates w.r.t. the
    gathered copy): wrt is turned into an autograd leaf and the cost graph re-evaluated;
  * x[idx] with ints, slices with symbolic bounds, integer vectors/matrices, tuples; T.set_subtensor on such a
    node returns the whole base with the indexed part replaced;
  * Unique(False, False, False)(x): sorted unique of the flattened input (numpy.unique);
  * T.dot with numpy.dot semantics; T.sum of a Python list stacks first.
"""
import sys
import types

import numpy as np
import torch

import os

# POI_SHIM_FLOAT=32 evaluates the same graphs in float32 (the reference's `floatX=float32` setting) -- used only to measure
# how far float32 arithmetic moves the reference's own trajectories from the float64 vectors (tests/test_theano_shim.py)
FLOAT = torch.float32 if os.environ.get("POI_SHIM_FLOAT") == "32" else torch.float64
NP_FLOAT = np.float32 if FLOAT == torch.float32 else np.float64
INT = torch.int64


# --------------------------------------------------------------------------------------------------
# evaluation environment
# --------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, bind=None, givens=None, parent=None):
        self.bind = bind if bind is not None else {}        # placeholder -> tensor
        self.givens = givens if givens is not None else {}  # placeholder -> Var
        self.memo = {}
        self.parent = parent                                # scan inner environments chain to the outer one

    def lookup(self, v):
        e = self
        while e is not None:
            if id(v) in e.memo:
                return True, e.memo[id(v)]
            e = e.parent
        return False, None

    def find_bind(self, v):
        e = self
        while e is not None:
            if v in e.bind:
                return True, e.bind[v]
            e = e.parent
        return False, None

    def find_given(self, v):
        e = self
        while e is not None:
            if v in e.givens:
                return e.givens[v]
            e = e.parent
        return None

    def root(self):
        e = self
        while e.parent is not None:
            e = e.parent
        return e


def as_var(x):
    if isinstance(x, Var):
        return x
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, Var) for e in x):
            return Stack([as_var(e) for e in x])
        return Const(np.asarray(x))
    return Const(x)


def to_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind in "iub":
        return torch.as_tensor(a.astype(np.int64))
    return torch.as_tensor(a.astype(NP_FLOAT))


def ev(v, env):
    if not isinstance(v, Var):
        return to_tensor(v)
    hit, val = env.lookup(v)
    if hit:
        return val
    val = v.compute(env)
    env.memo[id(v)] = val
    return val


# --------------------------------------------------------------------------------------------------
# nodes
# --------------------------------------------------------------------------------------------------
class Var:
    ndim = None
    __array_priority__ = 1000
    __hash__ = object.__hash__

    def compute(self, env):
        raise NotImplementedError

    # --- arithmetic ---
    def _bin(self, other, f, swap=False):
        o = as_var(other)
        return Op((o, self) if swap else (self, o), f)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, _true_div)
    def __rtruediv__(self, o): return self._bin(o, _true_div, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __pow__(self, o): return self._bin(o, _pow)
    def __rpow__(self, o): return self._bin(o, _pow, True)
    def __neg__(self): return Op((self,), lambda a: -a)
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)

    # --- structure ---
    def __getitem__(self, idx):
        return Subtensor(self, idx)

    def __iter__(self):
        raise TypeError("symbolic variable is not iterable")

    @property
    def T(self):
        return Op((self,), lambda a: a.t() if a.dim() == 2 else a.permute(*reversed(range(a.dim()))))

    @property
    def shape(self):
        return Shape(self)

    def sum(self, axis=None, keepdims=False):
        return t_sum(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, keepdims=False):
        return t_max(self, axis=axis, keepdims=keepdims)

    def reshape(self, shape, ndim=None):
        return t_reshape(self, shape)

    def dimshuffle(self, *pattern):
        if len(pattern) == 1 and isinstance(pattern[0], (list, tuple)):
            pattern = tuple(pattern[0])

        def f(a):
            perm = [p for p in pattern if p != 'x']
            out = a.permute(*perm) if perm else a
            for pos, p in enumerate(pattern):
                if p == 'x':
                    out = out.unsqueeze(pos)
            return out
        return Op((self,), f)

    def flatten(self):
        return Op((self,), lambda a: a.reshape(-1))

    def eval(self, inputs_to_values=None):
        env = Env(bind={k: to_tensor(v) for k, v in (inputs_to_values or {}).items()})
        return _to_numpy(ev(self, env))

    # numpy ufuncs applied to symbolic variables (PRME.py:121 calls Load_Data_prme.cal_dis, which is written with
    # np.sin / np.cos / np.arcsin / np.sqrt / np.power / np.multiply, on Theano variables)
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        f = _UFUNCS.get(ufunc.__name__)
        if method != "__call__" or f is None or kw:
            return NotImplemented
        return Op(tuple(as_var(i) for i in inputs), lambda *a: f(*[x.to(FLOAT) if not x.is_floating_point() else x for x in a]))


_UFUNCS = {"multiply": torch.mul, "add": torch.add, "subtract": torch.sub, "true_divide": torch.div, "divide": torch.div,
           "power": torch.pow, "sin": torch.sin, "cos": torch.cos, "arcsin": torch.asin, "sqrt": torch.sqrt, "exp": torch.exp,
           "log": torch.log, "negative": torch.neg, "absolute": torch.abs}


def _true_div(a, b):
    if not a.is_floating_point() and not b.is_floating_point():
        return a.to(FLOAT) / b.to(FLOAT)
    return a / b


def _pow(a, b):
    return torch.pow(a, b)


def _to_numpy(t):
    a = t.detach().cpu().numpy()
    return a.copy()


class Const(Var):
    def __init__(self, value):
        self.value = to_tensor(value)
        self.ndim = self.value.dim()

    def compute(self, env):
        return self.value


class Placeholder(Var):
    def __init__(self, kind, ndim, name=None):
        self.kind, self.ndim, self.name = kind, ndim, name

    def compute(self, env):
        hit, val = env.find_bind(self)
        if hit:
            return val
        g = env.find_given(self)
        if g is not None:
            return ev(g, env)
        raise KeyError("unbound symbolic input %r" % (self.name,))

    def cast(self, x):
        t = to_tensor(x)
        t = t.to(INT) if self.kind == 'i' else t.to(FLOAT)
        if t.dim() != self.ndim:
            raise TypeError("wrong number of dimensions: expected %d, got %d" % (self.ndim, t.dim()))
        return t


class Shared(Var):
    def __init__(self, value, name=None, borrow=False, **kw):
        self.tensor = to_tensor(value).clone()
        self.ndim = self.tensor.dim()
        self.name = name

    def compute(self, env):
        return self.tensor

    def get_value(self, borrow=False, return_internal_type=False):
        return _to_numpy(self.tensor)

    def set_value(self, value, borrow=False):
        self.tensor = to_tensor(value).clone()


class Op(Var):
    def __init__(self, args, f):
        self.args, self.f = tuple(args), f

    def compute(self, env):
        return self.f(*[ev(a, env) for a in self.args])


class Stack(Var):
    def __init__(self, items):
        self.items = items

    def compute(self, env):
        vals = [ev(a, env) for a in self.items]
        dt = FLOAT if any(v.is_floating_point() for v in vals) else INT
        return torch.stack([v.to(dt) for v in vals])


def _resolve_index(idx, env):
    def one(i):
        if isinstance(i, slice):
            def b(x):
                if x is None:
                    return None
                if isinstance(x, Var):
                    return int(ev(x, env).item())
                return int(x)
            return slice(b(i.start), b(i.stop), b(i.step))
        if isinstance(i, Var):
            t = ev(i, env)
            return int(t.item()) if t.dim() == 0 else t.to(INT)
        if isinstance(i, (list, np.ndarray)):
            return torch.as_tensor(np.asarray(i).astype(np.int64))
        if isinstance(i, (int, np.integer)):
            return int(i)
        if i is None or i is Ellipsis:
            return i
        raise TypeError("unsupported index %r" % (i,))
    if isinstance(idx, tuple):
        return tuple(one(i) for i in idx)
    return one(idx)


class Subtensor(Var):
    def __init__(self, base, idx):
        self.base, self.idx = base, idx
        if base.ndim is not None and not isinstance(idx, tuple):
            if isinstance(idx, (int, np.integer)):
                self.ndim = base.ndim - 1
            elif isinstance(idx, slice):
                self.ndim = base.ndim
            elif isinstance(idx, Var) and idx.ndim is not None:
                self.ndim = base.ndim - 1 + idx.ndim
            elif isinstance(idx, (list, np.ndarray)):
                self.ndim = base.ndim - 1 + np.ndim(idx)
        elif base.ndim is not None:
            n_int, adv = 0, []
            for i in idx:
                if isinstance(i, (int, np.integer)):
                    n_int += 1
                elif isinstance(i, Var):
                    adv.append(i.ndim)
                elif isinstance(i, (list, np.ndarray)):
                    adv.append(np.ndim(i))
            if all(a is not None for a in adv):
                self.ndim = base.ndim - n_int - len(adv) + (max(adv) if adv else 0)

    def compute(self, env):
        return ev(self.base, env)[_resolve_index(self.idx, env)]


class SetSubtensor(Var):
    def __init__(self, sub, value, inc=False):
        if not isinstance(sub, Subtensor):
            raise TypeError("set_subtensor needs x[idx] as its first argument")
        self.sub, self.value, self.inc = sub, as_var(value), inc
        self.ndim = sub.base.ndim

    def compute(self, env):
        base = ev(self.sub.base, env).clone()
        idx = _resolve_index(self.sub.idx, env)
        val = ev(self.value, env)
        if self.inc:
            base.index_put_((idx,) if not isinstance(idx, tuple) else idx, val.to(base.dtype), accumulate=True)
        else:
            base[idx] = val.to(base.dtype)
        return base


class Shape(Var):
    def __init__(self, x):
        self.x = x
        self.ndim = 1

    def compute(self, env):
        return torch.as_tensor(list(ev(self.x, env).shape), dtype=INT)

    def __iter__(self):
        if self.x.ndim is None:
            raise TypeError("shape of a variable with unknown ndim cannot be unpacked")
        return iter([self[i] for i in range(self.x.ndim)])

    def __len__(self):
        if self.x.ndim is None:
            raise TypeError("unknown ndim")
        return self.x.ndim


# --------------------------------------------------------------------------------------------------
# theano.tensor functions
# --------------------------------------------------------------------------------------------------
def _axis_arg(axis):
    return axis if axis is None or isinstance(axis, int) else tuple(axis)


def t_sum(x, axis=None, keepdims=False, **kw):
    x = as_var(x)
    ax = _axis_arg(axis)

    def f(a):
        if not a.is_floating_point() and a.dtype != INT:
            a = a.to(INT)
        return a.sum() if ax is None else a.sum(dim=ax, keepdim=keepdims)
    out = Op((x,), f)
    return out


def t_max(x, axis=None, keepdims=False):
    x = as_var(x)
    ax = _axis_arg(axis)
    return Op((x,), lambda a: a.max() if ax is None else a.amax(dim=ax, keepdim=keepdims))


def t_dot(a, b):
    def f(x, y):
        dt = FLOAT if (x.is_floating_point() or y.is_floating_point()) else INT
        x, y = x.to(dt), y.to(dt)
        if y.dim() <= 2:
            return torch.matmul(x, y)         # == numpy.dot for a second operand of rank <= 2
        return torch.tensordot(x, y, dims=([x.dim() - 1], [y.dim() - 2]))
    return Op((as_var(a), as_var(b)), f)


def t_concatenate(items, axis=0):
    items = [as_var(i) for i in items]

    def f(*vals):
        dt = FLOAT if any(v.is_floating_point() for v in vals) else INT
        return torch.cat([v.to(dt) for v in vals], dim=axis)
    out = Op(items, f)
    out.ndim = next((i.ndim for i in items if i.ndim is not None), None)
    return out


def _shape_list(shape, env):
    out = []
    for s in shape:
        out.append(int(ev(s, env).item()) if isinstance(s, Var) else int(s))
    return out


class Reshape(Var):
    def __init__(self, x, shape):
        self.x = as_var(x)
        self.shp = shape
        if isinstance(shape, (tuple, list)):
            self.ndim = len(shape)

    def compute(self, env):
        a = ev(self.x, env)
        if isinstance(self.shp, Var):
            shp = [int(v) for v in ev(self.shp, env).tolist()]
        else:
            shp = _shape_list(self.shp, env)
        return a.reshape(shp)


def t_reshape(x, shape, ndim=None):
    return Reshape(x, shape)


class Alloc(Var):
    def __init__(self, value, *shape):
        self.value, self.shp = as_var(value), shape
        self.ndim = len(shape)

    def compute(self, env):
        v = ev(self.value, env)
        shp = _shape_list(self.shp, env)
        return v.expand(*shp).clone() if shp else v.clone()


def t_ones_like(x):
    return Op((as_var(x),), lambda a: torch.ones_like(a))


def t_zeros_like(x):
    return Op((as_var(x),), lambda a: torch.zeros_like(a))


def t_arange(*args):
    vs = [as_var(a) for a in args]
    return Op(vs, lambda *a: torch.arange(*[int(x.item()) for x in a], dtype=INT))


def _unary(f):
    return lambda x: Op((as_var(x),), lambda a: f(a.to(FLOAT) if not a.is_floating_point() else a))


t_log, t_exp, t_sqrt, t_tanh = _unary(torch.log), _unary(torch.exp), _unary(torch.sqrt), _unary(torch.tanh)
t_sigmoid = _unary(torch.sigmoid)
t_abs = _unary(torch.abs)


def t_pow(a, b):
    return as_var(a) ** b


def t_gt(a, b):
    return as_var(a) > b


def t_lt(a, b):
    return as_var(a) < b


def t_set_subtensor(x, y, inplace=False, tolerate_inplace_aliasing=False):
    return SetSubtensor(x, y, inc=False)


def t_inc_subtensor(x, y, inplace=False, set_instead_of_inc=False, tolerate_inplace_aliasing=False):
    return SetSubtensor(x, y, inc=not set_instead_of_inc)


def t_nnet_softmax(x):
    return Op((as_var(x),), lambda a: torch.softmax(a, dim=-1))


class Rebroadcast:
    def __init__(self, *axis):
        pass

    def __call__(self, x):
        return x


def _placeholder_factory(kind, ndim):
    def make(name=None):
        return Placeholder(kind, ndim, name)
    return make


class Unique:
    """theano.tensor.extra_ops.Unique(return_index, return_inverse, return_counts): numpy.unique of the input."""

    def __init__(self, return_index=False, return_inverse=False, return_counts=False, axis=None):
        if return_index or return_inverse or return_counts:
            raise NotImplementedError("only Unique(False, False, False) is used by the reference")

    def __call__(self, x):
        out = Op((as_var(x),), lambda a: torch.unique(a.reshape(-1), sorted=True))
        out.ndim = 1
        return out


class IfElse(Var):
    def __init__(self, cond, a, b):
        self.cond, self.a, self.b = as_var(cond), as_var(a), as_var(b)

    def compute(self, env):
        return ev(self.a, env) if bool(ev(self.cond, env).item()) else ev(self.b, env)


def ifelse(cond, then_branch, else_branch, name=None):
    return IfElse(cond, then_branch, else_branch)


# --------------------------------------------------------------------------------------------------
# scan
# --------------------------------------------------------------------------------------------------
class Scan:
    def __init__(self, fn, sequences, outputs_info, non_sequences, n_steps):
        self.seqs = [as_var(s) for s in sequences]
        self.n_steps = n_steps
        self.seq_ph = [Placeholder('x', None, "scan_seq%d" % i) for i in range(len(self.seqs))]
        self.rec_init = [as_var(o) for o in outputs_info if o is not None]
        self.rec_ph = [Placeholder('x', None, "scan_prev%d" % i) for i in range(len(self.rec_init))]
        outs = fn(*(self.seq_ph + self.rec_ph + list(non_sequences)))
        if isinstance(outs, tuple) and len(outs) == 2 and isinstance(outs[1], dict):
            outs = outs[0]
        self.single = not isinstance(outs, (list, tuple))
        self.outs = [as_var(o) for o in ([outs] if self.single else outs)]
        if len(self.outs) != len(outputs_info):
            raise ValueError("scan: fn returned %d outputs for %d outputs_info" % (len(self.outs), len(outputs_info)))
        self.rec_pos = [k for k, o in enumerate(outputs_info) if o is not None]

    def run(self, env):
        key = ("scan", id(self))
        root = env
        hit, val = root.lookup(self)           # memo keyed on id(self)
        if hit:
            return val
        seq_vals = [ev(s, env) for s in self.seqs]
        if self.n_steps is not None:
            n = int(ev(self.n_steps, env).item()) if isinstance(self.n_steps, Var) else int(self.n_steps)
        else:
            n = min(int(s.shape[0]) for s in seq_vals)
        prev = [ev(i, env) for i in self.rec_init]
        collected = [[] for _ in self.outs]
        for t in range(n):
            inner = Env(parent=env)
            for ph, sv in zip(self.seq_ph, seq_vals):
                inner.bind[ph] = sv[t]
            for ph, pv in zip(self.rec_ph, prev):
                inner.bind[ph] = pv
            vals = [ev(o, inner) for o in self.outs]
            for k, v in enumerate(vals):
                collected[k].append(v)
            prev = [vals[k] for k in self.rec_pos]
        res = []
        for k, c in enumerate(collected):
            if c:
                res.append(torch.stack(c))
            else:
                res.append(torch.zeros((0,), dtype=FLOAT))
        env.memo[id(self)] = res
        return res


class ScanOut(Var):
    def __init__(self, scan, k):
        self.scan, self.k = scan, k

    def compute(self, env):
        return self.scan.run(env)[self.k]


def scan(fn, sequences=None, outputs_info=None, non_sequences=None, n_steps=None, truncate_gradient=-1,
         go_backwards=False, mode=None, name=None, profile=False, allow_gc=None, strict=False):
    if go_backwards:
        raise NotImplementedError
    sequences = [] if sequences is None else (list(sequences) if isinstance(sequences, (list, tuple)) else [sequences])
    single_info = outputs_info is not None and not isinstance(outputs_info, (list, tuple))
    infos = [outputs_info] if single_info else list(outputs_info or [])
    non_sequences = [] if non_sequences is None else (
        list(non_sequences) if isinstance(non_sequences, (list, tuple)) else [non_sequences])
    sc = Scan(fn, sequences, infos, non_sequences, n_steps)
    outs = [ScanOut(sc, k) for k in range(len(sc.outs))]
    return (outs[0] if (sc.single or single_info) else outs), {}


# --------------------------------------------------------------------------------------------------
# grad
# --------------------------------------------------------------------------------------------------
class GradSet:
    """d cost / d wrt[k] for a list of wrt nodes: each wrt value becomes an autograd leaf, the cost graph is
    re-evaluated on top of the leaves in a fresh environment (same inputs and givens)."""

    def __init__(self, cost, wrts):
        self.cost, self.wrts = cost, wrts

    def run(self, env):
        hit, val = env.lookup(self)
        if hit:
            return val
        root = env
        fresh = Env(bind=dict(root.bind), givens=dict(root.givens), parent=None)
        # carry the scan-level bindings too (grad inside scan is not used by the reference)
        e = root.parent
        while e is not None:
            for k, v in e.bind.items():
                fresh.bind.setdefault(k, v)
            for k, v in e.givens.items():
                fresh.givens.setdefault(k, v)
            e = e.parent
        leaves = []
        for w in self.wrts:
            with torch.no_grad():
                v = ev(w, Env(bind=fresh.bind, givens=fresh.givens)).detach().clone()
            if not v.is_floating_point():
                raise TypeError("grad with respect to an integer variable")
            v.requires_grad_(True)
            fresh.memo[id(w)] = v
            leaves.append(v)
        with torch.enable_grad():
            c = ev(self.cost, fresh)
            if c.dim() != 0:
                raise TypeError("cost must be a scalar")
            gs = torch.autograd.grad(c, leaves, allow_unused=True)
        out = [torch.zeros_like(l) if g is None else g.detach() for g, l in zip(gs, leaves)]
        env.memo[id(self)] = out
        return out


class GradOut(Var):
    def __init__(self, gset, k, ndim):
        self.gset, self.k, self.ndim = gset, k, ndim

    def compute(self, env):
        return self.gset.run(env)[self.k]


def grad(cost, wrt, consider_constant=None, disconnected_inputs='raise', **kw):
    single = not isinstance(wrt, (list, tuple))
    wrts = [wrt] if single else list(wrt)
    gs = GradSet(as_var(cost), wrts)
    outs = [GradOut(gs, k, w.ndim) for k, w in enumerate(wrts)]
    return outs[0] if single else outs


# --------------------------------------------------------------------------------------------------
# function
# --------------------------------------------------------------------------------------------------
class Function:
    def __init__(self, inputs, outputs=None, updates=None, givens=None, **kw):
        self.inputs = list(inputs)
        self.outputs = outputs
        if updates is None:
            updates = []
        self.updates = list(updates.items()) if isinstance(updates, dict) else list(updates)
        self.givens = dict(givens.items() if isinstance(givens, dict) else (givens or []))

    def __call__(self, *args):
        if len(args) != len(self.inputs):
            raise TypeError("expected %d arguments, got %d" % (len(self.inputs), len(args)))
        bind = {p: p.cast(a) for p, a in zip(self.inputs, args)}
        env = Env(bind=bind, givens={k: as_var(v) for k, v in self.givens.items()})
        with torch.no_grad():
            if self.outputs is None:
                outs = None
            elif isinstance(self.outputs, (list, tuple)):
                outs = [_to_numpy(ev(as_var(o), env)) for o in self.outputs]
            else:
                outs = _to_numpy(ev(as_var(self.outputs), env))
            new_vals = [(sh, ev(as_var(expr), env).detach().clone()) for sh, expr in self.updates]
        for sh, v in new_vals:
            if tuple(v.shape) != tuple(sh.tensor.shape):
                raise ValueError("update changes the shape of a shared variable")
            sh.tensor = v.to(sh.tensor.dtype)
        return outs


def function(inputs, outputs=None, updates=None, givens=None, **kw):
    return Function(inputs, outputs, updates, givens, **kw)


def shared(value, name=None, borrow=False, **kw):
    return Shared(value, name=name, borrow=borrow)


class RandomStreams:
    def __init__(self, seed=None, **kw):
        self.seed = seed

    def binomial(self, *a, **k):
        raise NotImplementedError("random streams are not on the in-scope train path")

    uniform = normal = binomial


# --------------------------------------------------------------------------------------------------
# module assembly
# --------------------------------------------------------------------------------------------------
def install():
    """Put `theano` and the sub-modules the reference imports into sys.modules.  Idempotent."""
    if "theano" in sys.modules and getattr(sys.modules["theano"], "__poi_shim__", False):
        return sys.modules["theano"]

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    th = mod("theano")
    th.__poi_shim__ = True
    th.config = types.SimpleNamespace(floatX="float32" if FLOAT == torch.float32 else "float64")
    th.shared, th.function, th.scan, th.grad = shared, function, scan, grad

    T = mod("theano.tensor")
    th.tensor = T
    for name, (kind, nd) in dict(iscalar=('i', 0), ivector=('i', 1), imatrix=('i', 2), lscalar=('i', 0), lvector=('i', 1),
                                 scalar=('f', 0), dscalar=('f', 0), fscalar=('f', 0), vector=('f', 1), fvector=('f', 1),
                                 dvector=('f', 1), matrix=('f', 2), fmatrix=('f', 2), dmatrix=('f', 2),
                                 tensor3=('f', 3), ftensor3=('f', 3), itensor3=('i', 3)).items():
        setattr(T, name, _placeholder_factory(kind, nd))
    T.sum, T.max, T.dot, T.concatenate, T.reshape, T.alloc = t_sum, t_max, t_dot, t_concatenate, t_reshape, Alloc
    T.ones_like, T.zeros_like, T.arange = t_ones_like, t_zeros_like, t_arange
    T.log, T.exp, T.sqrt, T.tanh, T.pow, T.gt, T.lt, T.abs_ = t_log, t_exp, t_sqrt, t_tanh, t_pow, t_gt, t_lt, t_abs
    T.grad, T.set_subtensor, T.inc_subtensor, T.Rebroadcast = grad, t_set_subtensor, t_inc_subtensor, Rebroadcast
    T.TensorVariable = Var

    nnet = mod("theano.tensor.nnet")
    T.nnet = nnet
    nnet.sigmoid, nnet.softmax = t_sigmoid, t_nnet_softmax
    nn2 = mod("theano.tensor.nnet.nnet")
    nnet.nnet = nn2
    nn2.softmax, nn2.sigmoid = t_nnet_softmax, t_sigmoid

    eo = mod("theano.tensor.extra_ops")
    T.extra_ops = eo
    eo.Unique = Unique

    rs = mod("theano.tensor.shared_randomstreams")
    T.shared_randomstreams = rs
    rs.RandomStreams = RandomStreams

    sb = mod("theano.sandbox")
    th.sandbox = sb
    mrg = mod("theano.sandbox.rng_mrg")
    sb.rng_mrg = mrg
    mrg.MRG_RandomStreams = RandomStreams

    ie = mod("theano.ifelse")
    th.ifelse = ie
    ie.ifelse = ifelse
    return th
