"""CPU oracle for the joint trainer's list scoring and token injection.  TEST INFRASTRUCTURE ONLY.

fp32 torch-CPU restatement, every function citing the reference lines it follows
(training/train_item_individual_token_joint.py under /root/reference).  Only tests/ may import it.

Parity status: PINNED - oracle/pin_joint_against_reference.py extracts the UNMODIFIED source of the reference's
`InfoNCELoss`, `MRREvaluator` and `MultiModalQwenEmbedding` classes (the file itself cannot be imported: it needs peft and
calls torch.cuda.set_device(0) at import, :33), executes them on seeded inputs, checks this restatement against them and
stores the reference's outputs in tests/golden/joint_scoring.npz.  `inject_tokens` is pinned against the unmodified
`MultiModalQwenEmbedding.forward` (:146-171) run with a stub tokenizer and an identity stand-in for the Qwen3 base model
(the injection loop touches neither): bit-exact.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F


def infonce_per_user(users, pos, negs, masks: Optional[torch.Tensor], temperature: float) -> torch.Tensor:
    """InfoNCELoss.forward before the final mean (:331-351)."""
    u = F.normalize(users.float(), p=2, dim=-1)                       # :333
    p = F.normalize(pos.float(), p=2, dim=-1)                         # :334
    n = F.normalize(negs.float(), p=2, dim=-1)                        # :335
    pos_sim = (u * p).sum(-1) / temperature                           # :336
    neg_sim = torch.bmm(u.unsqueeze(1), n.transpose(-2, -1)).squeeze(1) / temperature    # :337-340
    losses = []
    for i in range(u.shape[0]):                                       # :344-351
        valid = neg_sim[i][masks[i]] if masks is not None else neg_sim[i]
        all_sim = torch.cat([pos_sim[i:i + 1], valid])
        losses.append(-pos_sim[i] + torch.logsumexp(all_sim, dim=0))
    return torch.stack(losses)


def infonce_loss(users, pos, negs, masks=None, temperature: float = 0.07) -> torch.Tensor:
    return infonce_per_user(users, pos, negs, masks, temperature).mean()      # :352


def list_similarities(users, pos, negs_list: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Per user: cosine similarities of [positive] + its negatives (:405-413)."""
    u = F.normalize(users.float(), p=2, dim=-1)
    p = F.normalize(pos.float(), p=2, dim=-1)
    out = []
    for i in range(u.shape[0]):
        n = F.normalize(negs_list[i].float(), p=2, dim=-1)
        out.append(torch.matmul(u[i], torch.cat([p[i].unsqueeze(0), n], 0).t()))
    return out


def reciprocal_ranks(users, pos, negs_list: Sequence[torch.Tensor]) -> List[float]:
    """MRREvaluator._compute_batch_mrr after the model call (:405-418)."""
    out = []
    for sims in list_similarities(users, pos, negs_list):
        order = torch.argsort(sims, descending=True)                  # :414
        rank = (order == 0).nonzero(as_tuple=True)[0].item() + 1      # :415
        out.append(1.0 / rank)
    return out


def inject_tokens(text_embeds, input_ids, token_ids, history_item_query_tokens):
    """The placeholder overwrite of JointQwen3WithQFormer.forward (:160-171); token_ids [num_hist, Q]."""
    out = text_embeds.clone()
    B, nh, Q, _ = history_item_query_tokens.shape
    for i in range(nh):
        for j in range(Q):
            tid = int(token_ids[i, j])
            emb = history_item_query_tokens[:, i, j, :]
            for b in range(B):
                positions = (input_ids[b] == tid).nonzero(as_tuple=True)[0]
                if len(positions) > 0:
                    out[b, positions] = emb[b].to(out.dtype)
    return out
