"""Host-side algebra of the LayerNorm fold (unirec_b200.ops.fold_layernorm_weights; models/qformer.py:285-289, 371-375 followed
by the next nn.Linear): pure torch, runs without a GPU.  The CUDA epilogue evaluates exactly the right-hand sides below from the
row statistics (sum, sum of squares) that the producer GEMM wrote (unirec_linear_ln_bf16)."""
import torch

from unirec_b200 import ops


def _stats(x):
    # what the producer epilogue stores per row: partial (sum, sum of squares) per 128-column piece, added in index order
    M, H = x.shape
    parts = x.view(M, H // 128, 128)
    return torch.stack([parts.sum(2), (parts * parts).sum(2)], dim=-1)          # [M, parts, 2]


def test_linear_of_layernorm_equals_the_folded_form():
    g = torch.Generator().manual_seed(0)
    M, H, N, eps = 37, 256, 96, 1e-12
    x = torch.randn(M, H, generator=g, dtype=torch.float64) * 1.7 + 0.4          # non-zero row means
    w = torch.randn(N, H, generator=g, dtype=torch.float64) * 0.05
    b = torch.randn(N, generator=g, dtype=torch.float64)
    gamma = 1.0 + 0.1 * torch.randn(H, generator=g, dtype=torch.float64)
    beta = 0.1 * torch.randn(H, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.layer_norm(x, (H,), gamma, beta, eps) @ w.t() + b
    wf, bf, cf = ops.fold_layernorm_weights(w, b, gamma, beta)                   # W' is rounded to bf16 (tensor-core operand)
    assert wf.dtype == torch.bfloat16 and bf.dtype == torch.float32 and cf.dtype == torch.float32
    st = _stats(x).sum(1)
    mu = st[:, 0] / H
    rstd = torch.rsqrt((st[:, 1] / H - mu * mu).clamp_min(0) + eps)
    folded = rstd[:, None] * (x @ wf.double().t()) - (mu * rstd)[:, None] * cf.double()[None, :] + bf.double()[None, :]
    # the only difference is the bf16 rounding of W' = W o gamma (2^-9 relative per weight), not the algebra
    exact_w = (w * gamma[None, :])
    folded_exact = rstd[:, None] * (x @ exact_w.t()) - (mu * rstd)[:, None] * exact_w.sum(1)[None, :] + (b + w @ beta)[None, :]
    torch.testing.assert_close(folded_exact, ref, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(folded, ref, rtol=0, atol=0.05)
    # c is the row sum of the ROUNDED weight: the mean term cancels what the tensor cores accumulate, exactly
    torch.testing.assert_close(cf.double(), wf.double().sum(1), rtol=1e-6, atol=1e-6)


def test_layernorm_as_residual_form_and_partial_statistics():
    g = torch.Generator().manual_seed(1)
    M, H, eps = 19, 512, 1e-12
    x = (torch.randn(M, H, generator=g) * 2.0 - 0.3).to(torch.bfloat16).double()   # the stored bf16 values
    gamma = 1.0 + 0.1 * torch.randn(H, generator=g, dtype=torch.float64)
    beta = 0.1 * torch.randn(H, generator=g, dtype=torch.float64)
    st = _stats(x)
    assert tuple(st.shape) == (M, H // 128, 2)
    s = st.sum(1)
    mu = s[:, 0] / H
    rstd = torch.rsqrt((s[:, 1] / H - mu * mu).clamp_min(0) + eps)
    res = x * (rstd[:, None] * gamma[None, :]) + ((-mu * rstd)[:, None] * gamma[None, :] + beta[None, :])
    torch.testing.assert_close(res, torch.nn.functional.layer_norm(x, (H,), gamma, beta, eps), rtol=1e-9, atol=1e-9)


def test_fold_without_linear_bias():
    w = torch.randn(8, 128)
    gamma, beta = torch.ones(128), torch.full((128,), 0.5)
    wf, bf, cf = ops.fold_layernorm_weights(w, None, gamma, beta)
    torch.testing.assert_close(bf, w @ beta)
    torch.testing.assert_close(cf, wf.float().sum(1))
