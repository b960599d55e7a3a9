"""Reconstruction-quality evaluation of the item Q-Former (evaluation/evaluate_item_qformer.py:29-103) on the GPU path.

`evaluate_reconstruction` is the loop of the reference's `evaluate_reconstruction_quality` (:66-95) with its arithmetic
in one kernel per batch (`ops.reconstruction_metrics`) and a single device-to-host read at the end instead of three
`.item()` synchronisations per batch.  Same result keys and the same averaging conventions:
  val_recon_loss        = mean over BATCHES of (masked squared-error sum / number of valid fields of the batch)  (:74-76,:92)
  avg_cosine_similarity = mean over all valid FIELDS of cosine(reconstructed, original)                          (:79-93)
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


@torch.no_grad()
def evaluate_reconstruction(model, val_embeddings: torch.Tensor, val_masks: torch.Tensor, batch_size: int = 256,
                            device: Optional[torch.device] = None) -> Dict[str, float]:
    """val_embeddings [N, F, E] fp32 and val_masks [N, F] (the cached tensors of models/qformer_utils.py:121-147; host or
    device).  The model must be a `unirec_b200.modules.QFormerForItemRepresentation` in eval() on a CUDA device."""
    dev = device or next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("evaluate_reconstruction: the model must live on a CUDA device (no CPU path)")
    n = val_embeddings.shape[0]
    n_batches = 0
    per_batch = []                                      # device tensors: batch-wise [sq, cos, count]
    for lo in range(0, n, batch_size):
        x = val_embeddings[lo:lo + batch_size].to(dev, non_blocking=True).float()
        m = val_masks[lo:lo + batch_size].to(dev, non_blocking=True)
        out = model(x, m)
        per_batch.append(ops.reconstruction_metrics(out["reconstructed_fields"], x, m))
        n_batches += 1
    if n_batches == 0:
        return {"val_recon_loss": 0, "avg_cosine_similarity": 0}
    acc = torch.stack(per_batch).cpu()                  # the only synchronisation
    counts = acc[:, 2]
    loss = float((acc[:, 0] / counts).mean())           # a batch without any valid field is NaN in the reference too
    total = float(counts.sum())
    return {"val_recon_loss": loss, "avg_cosine_similarity": float(acc[:, 1].sum() / total) if total > 0 else 0}
