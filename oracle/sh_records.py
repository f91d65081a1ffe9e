"""TEST INFRASTRUCTURE (oracle) — not imported by the product.

Restatement of the identity behind the planned "colour-record" feature exchange (DESIGN.md §6, round-2 plan):
the SH gradient the reference's preprocess backward produces for ONE view
(submodules/diff-gaussian-rasterization/cuda_rasterizer/backward.cu:20-139, computeColorFromSH backward)

    dL/dsh[k, c] = basis_k(dir) * dL/dRGB[c]      dir = normalize(xyz - campos),  dL/dRGB[c] = 0 where the
                                                    forward clamped channel c to 0 (forward.cu:66-70), and the
                                                    whole row is 0 for a culled Gaussian (radii == 0)

depends on the view only through the camera centre and 3 numbers per Gaussian.  So the gradient summed over N
views can be rebuilt anywhere from N records of (masked dL/dRGB, visible) per Gaussian plus xyz and the N camera
centres — 16 bytes per Gaussian and view instead of 4 * 3 * M.  `colour_records` packs what one view contributes,
`sh_grad_from_records` rebuilds the sum in view order with the basis written as in backward.cu:36-128.
"""
import numpy as np

SH_C0 = np.float32(0.28209479177387814)
SH_C1 = np.float32(0.4886025119029199)
SH_C2 = np.array([1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
                  0.5462742152960396], np.float32)
SH_C3 = np.array([-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                  -0.4570457994644658, 1.445305721320277, -0.5900435899266435], np.float32)


def colour_records(dL_dcolor, clamped, radii):
    """[P,4] float32: dL/dRGB with clamped channels zeroed, w = 1 for a visible Gaussian (radii > 0) else 0.
    dL_dcolor [P,3] is K7's output (backward.cu:413-586), clamped [P,3] bool/uint8 and radii [P] come from K1."""
    rec = np.zeros((dL_dcolor.shape[0], 4), np.float32)
    vis = np.asarray(radii) > 0
    rec[:, :3] = np.where(np.asarray(clamped).astype(bool), np.float32(0), np.asarray(dL_dcolor, np.float32))
    rec[~vis, :3] = 0
    rec[:, 3] = vis.astype(np.float32)
    return rec


def sh_basis(xyz, campos, degree):
    """[P,16] float32 basis values for dir = normalize(xyz - campos) (entries above the active degree are 0)."""
    d = np.asarray(xyz, np.float32) - np.asarray(campos, np.float32)[None, :]
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)).astype(np.float32)
    x, y, z = (d[:, 0] / ln).astype(np.float32), (d[:, 1] / ln).astype(np.float32), (d[:, 2] / ln).astype(np.float32)
    b = np.zeros((d.shape[0], 16), np.float32)
    b[:, 0] = SH_C0
    if degree > 0:
        b[:, 1], b[:, 2], b[:, 3] = -SH_C1 * y, SH_C1 * z, -SH_C1 * x
    if degree > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        b[:, 4], b[:, 5] = SH_C2[0] * xy, SH_C2[1] * yz
        b[:, 6] = SH_C2[2] * (np.float32(2) * zz - xx - yy)
        b[:, 7], b[:, 8] = SH_C2[3] * xz, SH_C2[4] * (xx - yy)
        if degree > 2:
            b[:, 9] = SH_C3[0] * y * (np.float32(3) * xx - yy)
            b[:, 10] = SH_C3[1] * xy * z
            b[:, 11] = SH_C3[2] * y * (np.float32(4) * zz - xx - yy)
            b[:, 12] = SH_C3[3] * z * (np.float32(2) * zz - np.float32(3) * xx - np.float32(3) * yy)
            b[:, 13] = SH_C3[4] * x * (np.float32(4) * zz - xx - yy)
            b[:, 14] = SH_C3[5] * z * (xx - yy)
            b[:, 15] = SH_C3[6] * x * (xx - np.float32(3) * yy)
    return b


def sh_grad_from_records(xyz, campos_list, records_list, degree, M):
    """Sum over views (in list order, fp32) of basis_k(dir_view) * record_view.rgb -> [P,M,3]."""
    P = np.asarray(xyz).shape[0]
    out = np.zeros((P, M, 3), np.float32)
    ncoef = min(M, (degree + 1) ** 2)
    for campos, rec in zip(campos_list, records_list):
        b = sh_basis(xyz, campos, degree)
        contrib = (b[:, :ncoef, None] * rec[:, None, :3]).astype(np.float32)
        contrib[rec[:, 3] == 0] = 0  # a culled Gaussian's direction may be degenerate; its row is exactly zero
        out[:, :ncoef] = (out[:, :ncoef] + contrib).astype(np.float32)
    return out
