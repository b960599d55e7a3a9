"""Shim that imports the UNMODIFIED reference - from /root/reference in the authoring container, or from the verbatim
copy oracle/_ref/ (oracle/build_ref.py; git-ignored, travels with the gpurun snapshot) on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by oracle/pin_*_against_reference.py to (a) pin the CPU restatement in
oracle/qformer_oracle.py against the reference's own code and (b) generate the golden vectors under tests/golden/, and
by bench.py's `--impl reference` arm and same-GPU eager comparator (oracle/reference_runner.py).  Nothing under
unirec_b200/ imports it.

The reference was written against transformers ~4.x (models/qformer.py:39-44);
this image has 5.5.0, so a few moved / removed helpers are patched in before the
import.  No reference file is modified.
"""
import sys
import types

import os

_REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference/models") else _REF_COPY


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "qformer.py"))


def install():
    import torch  # noqa: F401
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    from transformers.modeling_utils import PreTrainedModel

    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def _unused(*a, **k):  # head pruning is never reached on this path
            raise NotImplementedError
        mu.find_pruneable_heads_and_indices = _unused
    if not hasattr(PreTrainedModel, "get_head_mask"):
        PreTrainedModel.get_head_mask = (
            lambda self, hm, n, *_: [None] * n if hm is None else hm)

    _orig_init_weights = PreTrainedModel.init_weights

    def _init_weights(self):
        if not hasattr(self, "all_tied_weights_keys"):
            self.all_tied_weights_keys = {}
        return _orig_init_weights(self)

    if not getattr(PreTrainedModel.init_weights, "_unirec_shim", False):
        _init_weights._unirec_shim = True
        PreTrainedModel.init_weights = _init_weights

    if "sentence_transformers" not in sys.modules:
        st = types.ModuleType("sentence_transformers")
        st.SentenceTransformer = object
        sys.modules["sentence_transformers"] = st

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return (QFormerForItemRepresentation, UserQFormer, PositionalEncoding) from the reference."""
    install()
    from models.qformer_utils import QFormerForItemRepresentation
    from models.user_sequence_encoder import PositionalEncoding
    from training.user_qformer_training import UserQFormer
    return QFormerForItemRepresentation, UserQFormer, PositionalEncoding
