"""Weight preparation for the deploy-form kernels: BatchNorm folding and RepVGG re-parameterisation.

Reference: fuse_conv_and_bn / fuse_model (yolov6/utils/torch_utils.py:50-94) and
RepVGGBlock.get_equivalent_kernel_bias / switch_to_deploy (yolov6/layers/common.py:257-319).  The
reference does this by mutating modules in fp32; here it is a pure function from the train-form
state_dict to per-op (weight KRSC, bias) pairs, computed in fp64 so the only rounding left is the
final cast (the reference's own fold drifts its outputs by ~1e-4).
"""
import torch

BN_EPS = 1e-3  # torch_utils.py:41-43


def _bn_affine(sd, p):
    g, b = sd[p + ".weight"].double(), sd[p + ".bias"].double()
    m, v = sd[p + ".running_mean"].double(), sd[p + ".running_var"].double()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


def fold_conv_bn(sd, p):
    """conv (no bias) + BN -> (W' [Cout,Cin,k,k], b' [Cout]) in fp64."""
    w = sd[p + ".conv.weight"].double()
    scale, shift = _bn_affine(sd, p + ".bn")
    return w * scale.view(-1, 1, 1, 1), shift


def fold_op(sd, op):
    """Returns (weight fp64 [Cout,kh,kw,Cin] (KRSC), bias fp64 [Cout]); convT returns 4 KRSC 1x1 weights."""
    n = op.name
    if op.layout == "rep":
        k3, b3 = fold_conv_bn(sd, n + ".rbr_dense")
        k1, b1 = fold_conv_bn(sd, n + ".rbr_1x1")
        k = k3 + torch.nn.functional.pad(k1, [1, 1, 1, 1])
        b = b3 + b1
        if n + ".rbr_identity.weight" in sd:
            scale, shift = _bn_affine(sd, n + ".rbr_identity")
            idx = torch.arange(op.cin)
            k[idx, idx, 1, 1] += scale
            b = b + shift
        return k.permute(0, 2, 3, 1).contiguous(), b
    if op.layout == "cba":
        k, b = fold_conv_bn(sd, n + ".block")
        return k.permute(0, 2, 3, 1).contiguous(), b
    if op.layout == "plain":
        return sd[n + ".weight"].double().permute(0, 2, 3, 1).contiguous(), sd[n + ".bias"].double()
    if op.layout == "convT":
        w = sd[n + ".upsample_transpose.weight"].double()          # [Cin, Cout, 2, 2]
        quads = [w[:, :, dy, dx].t().contiguous().view(op.cout, 1, 1, op.cin) for dy in range(2) for dx in range(2)]
        return quads, sd[n + ".upsample_transpose.bias"].double()
    raise ValueError(op.layout)
