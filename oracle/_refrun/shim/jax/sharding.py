"""Device-mesh stand-in: a Mesh is just a (px, py) grid of fake devices; `_shard_map` runs the
per-shard function once per block and reassembles the global array (what shard_map means for the
out_specs the reference uses, distributed.py:17-34)."""
import numpy as _np

from . import numpy as jnp

_ACTIVE = [None]
_AXIS_INDEX = {}


class PartitionSpec(tuple):
    def __new__(cls, *names):
        return super().__new__(cls, names)


class Mesh:
    def __init__(self, devices, axis_names):
        self.devices = _np.asarray(devices)
        self.axis_names = tuple(axis_names)
        _ACTIVE[0] = self

    @property
    def empty(self):
        return self.devices.size == 0

    @property
    def shape(self):
        return dict(zip(self.axis_names, self.devices.shape))


AbstractMesh = Mesh


class NamedSharding:
    def __init__(self, mesh, spec):
        self.mesh, self.spec = mesh, spec
        _ACTIVE[0] = mesh


def make_mesh(pdims, axis_names=('x', 'y')):
    return Mesh(_np.arange(int(_np.prod(pdims))).reshape(pdims), axis_names)


def active_mesh():
    return _ACTIVE[0]


def clear_mesh():
    _ACTIVE[0] = None


def current_axis_index(name):
    return _AXIS_INDEX[name]


def _nblocks(mesh, spec, ndim):
    out = []
    for d in range(ndim):
        name = spec[d] if d < len(spec) else None
        out.append(1 if name is None else mesh.shape[name])
    return out


def split_blocks(x, mesh, spec):
    x = _np.asarray(jnp._raw(x))
    nb = _nblocks(mesh, spec, x.ndim)
    return nb, {ij: x[tuple(slice(i * (x.shape[d] // nb[d]), (i + 1) * (x.shape[d] // nb[d]))
                              for d, i in enumerate(ij))]
                for ij in _np.ndindex(*nb)}


def assemble_blocks(blocks, nb):
    def rec(prefix, d):
        if d == len(nb):
            return blocks[tuple(prefix)]
        return _np.concatenate([rec(prefix + [i], d + 1) for i in range(nb[d])], axis=d)
    return rec([], 0)


def _shard_map(f, mesh=None, in_specs=None, out_specs=None, check_vma=False, axis_names=frozenset(), **kw):
    def wrapped(*args):
        specs = in_specs
        if not isinstance(specs, tuple) or isinstance(specs, PartitionSpec):
            specs = (specs,) * len(args)      # one spec = pytree prefix for every argument
        names = mesh.axis_names
        sizes = mesh.devices.shape
        outs = {}
        for ij in _np.ndindex(*sizes):
            for n, i in zip(names, ij):
                _AXIS_INDEX[n] = i
            local = []
            for a, spec in zip(args, specs):
                if spec is None or len(spec) == 0 or not hasattr(a, 'ndim') or a.ndim == 0:
                    local.append(a)
                    continue
                nb, blocks = split_blocks(a, mesh, spec)
                key = tuple(ij[names.index(spec[d])] if d < len(spec) and spec[d] is not None else 0
                            for d in range(len(nb)))
                local.append(jnp._wrap(blocks[key].copy()))
            outs[ij] = _np.asarray(jnp._raw(f(*local)))
        nd = outs[(0,) * len(sizes)].ndim
        nb = _nblocks(mesh, out_specs, nd)
        keyed = {}
        for ij, v in outs.items():
            key = tuple(ij[names.index(out_specs[d])] if d < len(out_specs) and out_specs[d] is not None else 0
                        for d in range(nd))
            keyed[key] = v
        return jnp._wrap(assemble_blocks(keyed, nb))
    return wrapped
