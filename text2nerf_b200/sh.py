"""Real spherical-harmonic bases, mirror of the reference's ``models/sh.py`` surface
(eval_sh :34-85, eval_sh_bases :87-133).  Only degree 2 is reachable from the render path
(SHRender, tensorBase.py:29-33), where the fused appearance kernel evaluates it in registers;
these tensor versions serve callers that use the functions directly."""
import torch

_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)
_C4 = (2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
       -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761)


def _polys(deg, x, y, z):
    """List of (coefficient, polynomial) for every basis function up to `deg`."""
    out = [(_C0, None)]
    if deg > 0:
        out += [(-_C1, y), (_C1, z), (-_C1, x)]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [(_C2[0], xy), (_C2[1], yz), (_C2[2], 2.0 * zz - xx - yy), (_C2[3], xz), (_C2[4], xx - yy)]
        if deg > 2:
            out += [(_C3[0], y * (3 * xx - yy)), (_C3[1], xy * z), (_C3[2], y * (4 * zz - xx - yy)),
                    (_C3[3], z * (2 * zz - 3 * xx - 3 * yy)), (_C3[4], x * (4 * zz - xx - yy)),
                    (_C3[5], z * (xx - yy)), (_C3[6], x * (xx - 3 * yy))]
        if deg > 3:
            out += [(_C4[0], xy * (xx - yy)), (_C4[1], yz * (3 * xx - yy)), (_C4[2], xy * (7 * zz - 1)),
                    (_C4[3], yz * (7 * zz - 3)), (_C4[4], zz * (35 * zz - 30) + 3), (_C4[5], xz * (7 * zz - 3)),
                    (_C4[6], (xx - yy) * (7 * zz - 1)), (_C4[7], xz * (xx - 3 * yy)),
                    (_C4[8], xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return out


def eval_sh_bases(deg, dirs):
    """[..., (deg+1)^2] basis values at unit directions [..., 3]."""
    assert 0 <= deg <= 4
    x, y, z = dirs.unbind(-1)
    result = torch.empty((*dirs.shape[:-1], (deg + 1) ** 2), dtype=dirs.dtype, device=dirs.device)
    for j, (c, poly) in enumerate(_polys(deg, x, y, z)):
        result[..., j] = c if poly is None else c * poly
    return result


def eval_sh(deg, sh, dirs):
    """Sum_j basis_j(dirs) * sh[..., C, j]  ->  [..., C]."""
    assert 0 <= deg <= 4
    assert (deg + 1) ** 2 == sh.shape[-1]
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    total = None
    for j, (c, poly) in enumerate(_polys(deg, x, y, z)):
        term = c * sh[..., j] if poly is None else (c * poly) * sh[..., j]
        total = term if total is None else total + term
    return total
