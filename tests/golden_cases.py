"""Golden-vector case table shared by oracle/pin_against_reference.py (generator, container
only) and the tests (consumers, CPU box and GPU box).  Weights and inputs are regenerated from
unirec_b200.synth (pure function of key/shape/seed); only the REFERENCE's outputs are stored in
tests/golden/*.npz."""

_ITEM_FULL = dict(hidden=1024, layers=12, inter=4096, num_query=32, field_dim=1024, num_fields=14)
_ITEM_SMALL = dict(hidden=256, layers=4, inter=512, num_query=32, field_dim=256, num_fields=6)
_USER_FULL = dict(hidden=1024, layers=4, inter=4096, num_query=64, input_dim=1024, num_predict=32)
_USER_SMALL = dict(hidden=256, layers=2, inter=512, num_query=64, input_dim=256, num_predict=8)

ITEM_CASES = {
    # reference shapes, reference-like init; row 1 has Bernoulli field presence, row 2 has every
    # field masked (all-masked => uniform attention)
    "full": dict(model=_ITEM_FULL, heads=16, seed=0, attn_std=0.02,
                 input=dict(batch=3, num_fields=14, dim=1024, seed=1, presence=0.6, all_masked_row=2)),
    # sharper attention (larger q/k weights) so that softmax is far from uniform; over 12 layers a bf16
    # rounding that flips an attention winner is amplified, so this case is bounded by 1.5x the error
    # of the bf16 precision model on the same input (oracle/bf16_precision_model.py)
    "sharp": dict(model=_ITEM_FULL, heads=16, seed=3, attn_std=0.08,
                  input=dict(batch=2, num_fields=14, dim=1024, seed=4, presence=0.8)),
    # the same sharp weights, 2 layers (one with cross-attention): errors cannot compound, so the
    # standard tolerance applies even though softmax is near one-hot
    "sharp2": dict(model=dict(_ITEM_FULL, layers=2), heads=16, seed=3, attn_std=0.08,
                   input=dict(batch=2, num_fields=14, dim=1024, seed=4, presence=0.8)),
    # small generic-dims model: 4 heads x 64, 6 fields
    "small": dict(model=_ITEM_SMALL, heads=4, seed=5, attn_std=0.1,
                  input=dict(batch=5, num_fields=6, dim=256, seed=6, clip_field=2, presence=0.7,
                             all_masked_row=4)),
}

USER_CASES = {
    "full": dict(model=_USER_FULL, heads=16, seed=7, attn_std=0.02,
                 input=dict(batch=2, max_items=5, tokens_per_item=32, dim=1024, seed=8, ragged=True)),
    "long": dict(model=_USER_FULL, heads=16, seed=9, attn_std=0.06,
                 input=dict(batch=1, max_items=50, tokens_per_item=32, dim=1024, seed=10, ragged=False)),
    "small": dict(model=_USER_SMALL, heads=4, seed=11, attn_std=0.1,
                  input=dict(batch=4, max_items=7, tokens_per_item=32, dim=256, seed=12, ragged=True)),
}

SCORING_CASE = dict(users=8, cands=4096, dim=1024, k=100, seed=13)

# evaluation/evaluate_item_qformer.py:40-103 run on the 'small' item model over 23 synthetic items in batches of 8
# (the last batch is ragged); tests/golden/eval_metrics.npz holds the reference function's two result numbers
EVAL_CASE = dict(item_case="small", batch_size=8,
                 input=dict(batch=23, num_fields=6, dim=256, seed=77, clip_field=2, presence=0.7))
