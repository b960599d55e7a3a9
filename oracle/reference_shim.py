"""In-container shim that imports the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by oracle/pin_against_reference.py to (a) pin the
CPU restatement in oracle/qformer_oracle.py against the reference's own code and
(b) generate the golden vectors under tests/golden/.  /root/reference does not
exist on the GPU box, so nothing in tests/, bench.py or smoke() imports this file.

The reference was written against transformers ~4.x (models/qformer.py:39-44);
this image has 5.5.0, so a few moved / removed helpers are patched in before the
import.  No reference file is modified.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


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
