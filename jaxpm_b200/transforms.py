"""Forward mode and batching of the force-loop operators - the two JAX transformations the reference's tests use
besides `grad` (/root/reference/tests/test_distributed_pm.py:313-320 `jacfwd` / `jacrev` of pm_forces, :335-409
`vmap` over stacked initial conditions).

In a JAX deployment these are the `jvp` rule and the `vmap_method` of the FFI calls; here they are
plain functions over the same C-ABI entry points:

* JVP of paint with respect to positions = jpm_cic_paintgrad_f32 (the transpose of the readgrad kernel);
  JVP of read = readgrad . tangent + read(mesh tangent); pm_forces composes them with the (linear) k-space chain;
* batched call == loop of single calls over the leading axis (`vmap_method="sequential"`; every element is a
  full-GPU workload, so there is nothing to gain from interleaving them), with a batched C-ABI entry for the
  hot one (jpm_sim_forces_batched).
"""
import torch

from . import ops
from ._lib import as_f32, call, ptr, stream


def cic_paint_jvp(positions, tangent, mesh_shape, relative=False, weight=1.0):
    """d paint(positions)[tangent]: mesh tangent of the painted density for a position tangent."""
    pos = as_f32(positions)
    tan = as_f32(tangent, pos.device)
    mesh = torch.zeros(tuple(mesh_shape), dtype=torch.float32, device=pos.device)
    n = pos.numel() // 3
    wp, ws, keep = ops._wargs(weight, n, pos.device)
    call("jpm_cic_paintgrad_f32", stream(), ptr(mesh), ptr(pos), ptr(tan), wp, ws, n, *mesh.shape, 0, 0, int(relative))
    return mesh


def cic_read_jvp(mesh, positions, mesh_tangent, pos_tangent, relative=False):
    """(read, d read): tangent of read(mesh, positions) for tangents of both arguments (either may be None)."""
    m = as_f32(mesh)
    pos = as_f32(positions, m.device)
    val, grad = ops.cic_readgrad(m, pos, relative, want_value=True, want_grad=pos_tangent is not None)
    out = torch.zeros_like(val)
    if pos_tangent is not None:
        out = (grad * as_f32(pos_tangent, m.device)).sum(-1)
    if mesh_tangent is not None:
        rd = ops.cic_read_dx(as_f32(mesh_tangent, m.device), pos) if relative else ops.cic_read(mesh_tangent, pos)
        out = out + rd.reshape(out.shape)
    return val, out


def pm_forces_jvp(positions, tangent, mesh_shape=None, paint_absolute_pos=True, r_split=0.0):
    """(F, dF) = (pm_forces(x), d pm_forces(x)[v]) - forward mode through paint -> k-space chain -> read
    (jaxpm/pm.py:12-58).  One extra paint (gradient weights), one extra k-space chain, three readgrads."""
    pos = as_f32(positions)
    v = as_f32(tangent, pos.device)
    relative = not paint_absolute_pos
    mesh_shape = tuple(pos.shape[:3]) if (relative or mesh_shape is None) else tuple(mesh_shape)
    plan = ops.get_plan(mesh_shape, pos.device)
    rho = torch.zeros(plan.shape, dtype=torch.float32, device=pos.device)
    if relative:
        ops.cic_paint_dx_(rho, pos)
    else:
        ops.cic_paint_(rho, pos)
    f3 = ops.force_meshes_from_density(rho, plan, r_split)
    F = ops.cic_read3(f3, pos, 1.0, relative)
    df3 = ops.force_meshes_from_density(cic_paint_jvp(pos, v, mesh_shape, relative), plan, r_split)
    dF = ops.cic_read3(df3, pos, 1.0, relative)
    for c in range(3):
        _, g = ops.cic_readgrad(f3[c], pos, relative, want_value=False)
        dF[..., c] += (g * v).sum(-1).reshape(dF.shape[:-1])
    return F, dF


# ---- batching ---------------------------------------------------------------------------------------------
def pm_forces_batched(positions, mesh_shape=None, paint_absolute_pos=True, r_split=0.0):
    """positions [B, ..., 3] -> forces [B, ..., 3]; element b == pm_forces(positions[b]).  Power-of-two meshes run the
    batched C-ABI entry on one cached positions-only resident state."""
    pos = as_f32(positions)
    relative = not paint_absolute_pos
    shape = tuple(pos.shape[1:4]) if (relative or mesh_shape is None) else tuple(mesh_shape)
    if not ops.fast_path_shape(shape) or pos.shape[0] == 0:
        from .pm import pm_forces
        return torch.stack([pm_forces(pos[b], mesh_shape=shape, paint_absolute_pos=paint_absolute_pos, r_split=r_split)
                            for b in range(pos.shape[0])]) if pos.shape[0] else pos.clone()
    npart = pos[0].numel() // 3
    pshape = tuple(pos.shape[1:4]) if pos.dim() == 5 else (1, 1, npart)
    sim = ops.Sim(shape, pshape, relative, pos.device, tile=16 if min(shape) >= 64 else 8, margin=1, positions_only=True)
    out = torch.empty_like(pos)
    call("jpm_sim_forces_batched", sim.handle, stream(), ptr(pos), ptr(out), pos.shape[0], 1.0, float(r_split), None, 0, 0.0)
    return out


def cic_paint_dx_batched(displacements, weight=1.0):
    from .painting import cic_paint_dx
    return torch.stack([cic_paint_dx(d, weight=weight) for d in as_f32(displacements)])


def cic_read_dx_batched(meshes, displacements):
    from .painting import cic_read_dx
    return torch.stack([cic_read_dx(m, d) for m, d in zip(as_f32(meshes), as_f32(displacements))])


def lpt_batched(cosmo, initial_conditions, a=0.1, order=1):
    """tests/test_distributed_pm.py:388-409: vmap of lpt over stacked initial conditions (relative mode)."""
    from .pm import lpt
    outs = [lpt(cosmo, ic, a=a, order=order) for ic in as_f32(initial_conditions)]
    return tuple(torch.stack([o[i] for o in outs]) for i in range(3))
