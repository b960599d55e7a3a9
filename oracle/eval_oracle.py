"""CPU oracle of the reconstruction-quality evaluation.  TEST INFRASTRUCTURE ONLY.

fp32 torch-CPU restatement of evaluation/evaluate_item_qformer.py:66-95 (the loop of `evaluate_reconstruction_quality`),
with the model passed in as a callable.  Parity status: PINNED - oracle/pin_eval_against_reference.py runs the UNMODIFIED
reference function (shimmed reference model, synthetic weights, a temporary field-embedding cache) and stores its two
result numbers in tests/golden/eval_metrics.npz; tests/test_joint_cpu.py re-checks this restatement against them.
Only tests/ may import this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def batch_metrics(reconstructed, original, attention_mask):
    """(masked MSE of the batch, sum of cosine similarities over valid fields, number of valid fields) - :73-88."""
    rec, orig = reconstructed.float(), original.float()
    mask = attention_mask.to(rec.dtype)
    unreduced = F.mse_loss(rec, orig, reduction="none")                               # :74
    loss = (unreduced * mask.unsqueeze(-1)).sum() / mask.sum()                        # :75
    valid = attention_mask.bool()                                                     # :79
    o, r = orig[valid], rec[valid]                                                    # :81-82
    cos = torch.sum(F.normalize(o, p=2, dim=-1) * F.normalize(r, p=2, dim=-1), dim=-1)   # :85-88
    return float(loss), float(cos.sum()), int(o.shape[0])


def reconstruction_quality(model_fn, val_embeddings, val_masks, batch_size):
    """model_fn(fields, mask) -> reconstructed_fields.  Returns the reference's result dict (:92-103)."""
    total_loss, total_cos, n_valid, n_batches = 0.0, 0.0, 0, 0
    for lo in range(0, val_embeddings.shape[0], batch_size):
        x, m = val_embeddings[lo:lo + batch_size], val_masks[lo:lo + batch_size]
        loss, cos, n = batch_metrics(model_fn(x, m), x, m)
        total_loss += loss
        total_cos += cos
        n_valid += n
        n_batches += 1
    return {"val_recon_loss": total_loss / n_batches if n_batches else 0,
            "avg_cosine_similarity": total_cos / n_valid if n_valid else 0}
