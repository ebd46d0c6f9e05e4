"""CPU oracle of the hot path - TEST INFRASTRUCTURE ONLY.

A plain fp32 (fp64 where the reference uses numpy float64) CPU restatement of the
reference's algorithm for the path named in BASELINE.json: per-channel normalisation
(code/dataset.py, code/preprocess_helpers.py), the DWI / DCE encoders and the late-fusion
head (code/model_module.py, code/transformer_model.py).  Every function cites the
reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker or the timed CPU baseline - never as
part of the product path (the product fails loudly without its CUDA library).

Parity pinning: the reference ships no golden vectors or tests, so
this oracle is pinned against outputs of the UNMODIFIED reference modules imported from
/root/reference/code in the authoring container: oracle/make_golden.py generates
tests/golden/*.npz (inputs are regenerated from seeds, expected outputs are stored) and
tests/test_oracle_golden.py checks the restatement against them on CPU.
The timm ViT-B/16 backbone is third-party and absent (un-pinned dependency): that part is
"parity unpinned" and documented as such in DESIGN.md.
"""
