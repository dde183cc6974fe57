"""Row f2: the coefficient network's Dense layers as FP64 tensor-core GEMMs of the library (csrc/dense_gemm.cu) against the
plain torch float64 composite of the same layers (flax Dense / LayerNorm / ELU as grad_dft/functional.py:793-822 uses them)
and against the CPU oracle's dm21_mlp: values and every cotangent (input, kernels, biases, LayerNorm parameters)."""
import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200 import ops

pytestmark = pytest.mark.gpu
F64 = torch.float64
RTOL = 1e-11


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def composite_block(x, kernel, kb, scale, bias, eps=1e-6):
    z = x @ kernel + kb + x
    mu = z.mean(dim=-1, keepdim=True)
    var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)
    return torch.nn.functional.elu((z - mu) * torch.rsqrt(var + eps) * scale + bias)


@pytest.mark.parametrize("N,K,Wd", [(1000, 11, 256), (33, 256, 3), (1, 7, 5), (4097, 256, 256), (700, 40, 72), (129, 2, 8)])
def test_dense_layer(cuda_device, N, K, Wd):
    g = torch.Generator().manual_seed(N + K + Wd)
    x = torch.randn(N, K, generator=g, dtype=F64).to(cuda_device).requires_grad_(True)
    kernel = (torch.randn(K, Wd, generator=g, dtype=F64) / K ** 0.5).to(cuda_device).requires_grad_(True)
    bias = torch.randn(Wd, generator=g, dtype=F64).to(cuda_device).requires_grad_(True)
    cot = torch.randn(N, Wd, generator=g, dtype=F64).to(cuda_device)
    with ops.first_order_build():
        assert ops.dense_supported(x, kernel)
        y = ops.dense_layer(x, kernel, bias)
    y_ref = x @ kernel + bias
    assert rel(y, y_ref) < RTOL
    got = torch.autograd.grad(y, [x, kernel, bias], cot)
    ref = torch.autograd.grad(y_ref, [x, kernel, bias], cot)
    for a, b in zip(got, ref):
        assert a.shape == b.shape and rel(a, b) < RTOL
    # outside a first-order build (second-order requests) the layer stays the host-framework composite
    assert not ops.dense_supported(x, kernel)


@pytest.mark.parametrize("chain", [False, True])  # reverse pass: plain GEMM + streaming reverse | one GEMM with the reverse in its epilogue
@pytest.mark.parametrize("N,W,nb", [(1000, 256, 3), (37, 8, 2), (2048, 64, 1), (513, 200, 2), (31, 256, 6), (1, 16, 1)])
def test_residual_trunk(cuda_device, N, W, nb, chain, monkeypatch):
    monkeypatch.setattr(ops, "DENSE_BWD_CHAIN", chain)
    g = torch.Generator().manual_seed(N + W + nb)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64)  # noqa: E731
    x = rn(N, W).to(cuda_device).requires_grad_(True)
    blocks = []
    for _ in range(nb):
        blocks.append(tuple(t.to(cuda_device).requires_grad_(True) for t in
                            (torch.eye(W, dtype=F64) + rn(W, W) * (2.0 / W) ** 0.5, 0.1 * rn(W), 1.0 + 0.2 * rn(W), 0.3 * rn(W))))
    cot = rn(N, W).to(cuda_device)
    with ops.first_order_build():
        assert ops.residual_trunk_supported(x, W)
        y = ops.residual_trunk(x, blocks)
    y_ref = x
    for blk in blocks:
        y_ref = composite_block(y_ref, *blk)
    assert rel(y, y_ref) < RTOL
    leaves = [x] + [t for blk in blocks for t in blk]
    got = torch.autograd.grad(y, leaves, cot)
    ref = torch.autograd.grad(y_ref, leaves, cot)
    for k, (a, b) in enumerate(zip(got, ref)):
        assert a.shape == b.shape and rel(a, b) < 1e-10, (k, rel(a, b))
    # bitwise run-to-run reproducible (fixed-order split-K and column reductions)
    with ops.first_order_build():
        y2 = ops.residual_trunk(x, blocks)
    got2 = torch.autograd.grad(y2, leaves, cot)
    assert torch.equal(y, y2) and all(torch.equal(a, b) for a, b in zip(got, got2))


def test_dm21_network_matches_oracle(cuda_device):
    """DM21's default_nn through NeuralFunctional (first layer, six fused blocks, head) against oracle.dm21_mlp, values and
    parameter / input gradients (1e-9, the bar VERDICT names for the fused Dense layer)."""
    params = oracle.dm21_mlp_init(seed=11)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3000, 11, generator=g, dtype=F64) * 2.0 + 1e-3
    pl = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xl = x.clone().requires_grad_(True)
    ref = oracle.dm21_mlp(pl, xl)
    cot = torch.randn(ref.shape, generator=g, dtype=F64)
    ref_g = torch.autograd.grad(ref, [xl] + list(pl.values()), cot)
    fun = gd.DM21()
    pd = {k: v.to(cuda_device).requires_grad_(True) for k, v in params.items()}
    xd = x.to(cuda_device).requires_grad_(True)
    with ops.first_order_build():
        out = fun.apply(pd, xd)
    assert rel(out.cpu(), ref) < 1e-9
    got = torch.autograd.grad(out, [xd] + list(pd.values()), cot.to(cuda_device))
    for a, b, name in zip(got, ref_g, ["x"] + list(pd)):
        assert rel(a.cpu(), b) < 1e-9, name
