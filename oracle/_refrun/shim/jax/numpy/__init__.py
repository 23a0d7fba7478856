"""`jax.numpy` stand-in on NumPy with JAX's dtype behaviour (see oracle/_refrun/__init__.py).

An `Arr` is an ndarray subclass that (a) is immutable in the JAX sense — augmented assignment
rebinds instead of writing in place, updates go through `.at[...]`; (b) promotes like JAX:
Python scalars are weak, integer arrays never widen a float array, and with x64 disabled every
64-bit result is demoted to 32 bits.
"""
import numpy as _np

X64 = False
pi = _np.pi
newaxis = None
inf = _np.inf
float32, float64, int32, int64, complex64, complex128 = (_np.float32, _np.float64, _np.int32, _np.int64,
                                                         _np.complex64, _np.complex128)
ndarray = _np.ndarray

_DEMOTE = {_np.dtype('float64'): _np.float32, _np.dtype('int64'): _np.int32,
           _np.dtype('complex128'): _np.complex64, _np.dtype('uint64'): _np.uint32}


def _raw(x):
    return x.view(_np.ndarray) if isinstance(x, Arr) else x


def _wrap(x):
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, (_np.ndarray, _np.generic)):
        x = _np.asarray(x)
        if not X64 and x.dtype in _DEMOTE:
            x = x.astype(_DEMOTE[x.dtype])
        return x.view(Arr)
    return x


def _promote(inputs):
    """JAX-style promotion of ufunc inputs: Python scalars stay weak; among arrays the widest
    float/complex wins and integer arrays are cast to it (numpy would go to float64)."""
    arrs = [i for i in inputs if isinstance(i, (_np.ndarray, _np.generic))]
    inexact = [a.dtype for a in arrs if a.dtype.kind in 'fc']
    has_pyfloat = any(isinstance(i, (float, complex)) and not isinstance(i, _np.generic) for i in inputs)
    if inexact:
        tgt = _np.result_type(*inexact)
    elif has_pyfloat and arrs:
        tgt = _np.dtype(_np.float64 if X64 else _np.float32)
        if any(isinstance(i, complex) for i in inputs):
            tgt = _np.result_type(tgt, _np.complex64)
    else:
        return inputs
    out = []
    for i in inputs:
        if isinstance(i, (_np.ndarray, _np.generic)) and i.dtype.kind in 'iub':
            i = _np.asarray(i).astype(tgt if tgt.kind == 'f' else _np.result_type(tgt).type(0).real.dtype)
        out.append(i)
    return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


def _norm_index(arr, idx):
    """Advanced integer index tuple -> (flat-safe index tuple, in-bounds mask) with JAX semantics:
    negative indices wrap once, out-of-range ones are flagged (scatter drops them, gather fills)."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    if all(isinstance(i, (slice, int, type(None), type(Ellipsis))) for i in idx):
        return idx, None
    ok = True
    fixed = []
    for ax, i in enumerate(idx):
        i = _np.asarray(_raw(i))
        n = arr.shape[ax]
        i = _np.where(i < 0, i + n, i)
        ok = ok & (i >= 0) & (i < n)
        fixed.append(i)
    fixed = tuple(_np.where(ok, i, 0) for i in fixed)
    return fixed, _np.broadcast_to(ok, fixed[0].shape)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def _apply(self, vals, op):
        out = _np.array(_raw(self.arr), copy=True)
        idx, ok = _norm_index(out, self.idx)
        vals = _raw(vals)
        if ok is None:
            if out[idx].size == 0:       # jax returns x unchanged for an empty slice shape
                return _wrap(out)
            if op == 'add':
                out[idx] = out[idx] + _np.asarray(vals, dtype=out.dtype)
            else:
                out[idx] = vals
            return _wrap(out)
        vals = _np.broadcast_to(_np.asarray(vals, dtype=out.dtype), ok.shape)
        sel = tuple(i[ok] for i in idx)
        if op == 'add':
            _np.add.at(out, sel, vals[ok])      # unordered in XLA; sequential fp32 here
        else:
            out[sel] = vals[ok]
        return _wrap(out)

    def add(self, vals, **kw):
        return self._apply(vals, 'add')

    def set(self, vals, **kw):
        return self._apply(vals, 'set')

    def get(self, mode=None, fill_value=None, **kw):
        a = _raw(self.arr)
        idx, ok = _norm_index(a, self.idx)
        v = a[idx]
        if ok is not None:
            if mode in ('drop', 'fill'):
                v = _np.where(ok.reshape(ok.shape + (1,) * (v.ndim - ok.ndim)), v,
                              0 if fill_value is None else fill_value).astype(a.dtype)
        return _wrap(v)


class Arr(_np.ndarray):
    __array_priority__ = 1000

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        ins = _promote([_raw(i) for i in inputs])
        if out is not None:
            kw['out'] = tuple(_raw(o) for o in out)
        return _wrap(getattr(ufunc, method)(*ins, **kw))

    def __array_function__(self, func, types, args, kwargs):
        def strip(x):
            if isinstance(x, Arr):
                return _raw(x)
            if isinstance(x, (list, tuple)):
                return type(x)(strip(v) for v in x)
            return x
        return _wrap(func(*strip(args), **{k: strip(v) for k, v in kwargs.items()}))

    @property
    def at(self):
        return _At(self)

    def astype(self, dt, **kw):
        return _wrap(_raw(self).astype(dt, **kw))

    def flatten(self):
        return self.reshape(-1)

    def __setitem__(self, k, v):
        raise TypeError("JAX arrays are immutable (use .at[...])")

    def _new(op):
        def f(self, o):
            return getattr(self, op)(o)
        return f
    __iadd__, __isub__, __imul__ = _new('__add__'), _new('__sub__'), _new('__mul__')
    __itruediv__, __ifloordiv__, __imod__ = _new('__truediv__'), _new('__floordiv__'), _new('__mod__')
    __iand__, __ior__, __ipow__ = _new('__and__'), _new('__or__'), _new('__pow__')
    del _new


def _strip(x):
    if isinstance(x, Arr):
        return _raw(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_strip(v) for v in x)
    return x


def asarray(x, dtype=None):
    if isinstance(x, Arr) and dtype is None:
        return x
    return _wrap(_np.asarray(_strip(x), dtype=dtype))


def array(x, dtype=None):
    return _wrap(_np.array(_strip(x), dtype=dtype))


def zeros(shape, dtype=None, device=None):
    return _wrap(_np.zeros(shape, dtype=dtype or (_np.float64 if X64 else _np.float32)))


def ones(shape, dtype=None, device=None):
    return _wrap(_np.ones(shape, dtype=dtype or (_np.float64 if X64 else _np.float32)))


def empty(shape, dtype=None, device=None):
    return zeros(shape, dtype)


def full(shape, fill_value, dtype=None):
    return _wrap(_np.full(shape, _raw(fill_value), dtype=dtype))


def zeros_like(x, dtype=None, shape=None):
    x = _np.asarray(_raw(x))
    return _wrap(_np.zeros(x.shape if shape is None else shape, dtype=dtype or x.dtype))


def isscalar(x):
    return _np.isscalar(x) or (hasattr(x, 'ndim') and x.ndim == 0)


def isrealobj(x):
    return _np.isrealobj(_raw(x))


def expand_dims(x, axis):
    return _wrap(_np.expand_dims(_np.asarray(_raw(x)), axis))


def bincount(x, weights=None, minlength=0, length=None):
    n = length if length is not None else minlength
    w = None if weights is None else _np.asarray(_raw(weights))
    return _wrap(_np.bincount(_np.asarray(_raw(x)), weights=w, minlength=n)[:n or None]
                 .astype(w.dtype if w is not None else _np.int32))


def interp(x, xp, fp):
    return _wrap(_np.interp(_np.asarray(_raw(x)), _np.asarray(_raw(xp)), _np.asarray(_raw(fp))))


class _FFT:
    @staticmethod
    def fftn(x, norm=None):
        import scipy.fft as sf
        return _wrap(sf.fftn(_np.asarray(_raw(x)), norm=norm))

    @staticmethod
    def ifftn(x, norm=None):
        import scipy.fft as sf
        return _wrap(sf.ifftn(_np.asarray(_raw(x)), norm=norm))


fft = _FFT()


def __getattr__(name):
    """Everything else: the NumPy function of the same name, results wrapped as `Arr`
    (ufuncs go through Arr.__array_ufunc__ so that JAX promotion applies)."""
    f = getattr(_np, name)
    if isinstance(f, _np.ufunc):
        def uf(*a, **k):
            a = tuple(x if isinstance(x, Arr) else (_wrap(_np.asarray(x)) if isinstance(x, (_np.ndarray, _np.generic, list, tuple)) else x) for x in a)
            if not any(isinstance(x, Arr) for x in a):
                a = (_wrap(_np.asarray(a[0])),) + a[1:]
            return f(*a, **k)
        return uf
    if callable(f):
        def fn(*a, **k):
            def strip(x):
                if isinstance(x, Arr):
                    return _raw(x)
                if isinstance(x, (list, tuple)):
                    return type(x)(strip(v) for v in x)
                return x
            return _wrap(f(*strip(a), **{kk: strip(v) for kk, v in k.items()}))
        return fn
    return f
