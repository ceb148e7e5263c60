"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch fp32 / numpy) of the reference's algorithm for the DiffewS hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only as the
checker.  The product package `diffews_b200` never imports `oracle`.

PARITY UNPINNED for the UNet / VAE / scheduler part; PINNED for oracle/data.py (unmodified reference datasets) and for
oracle/metric.py classify_prediction + AverageMeter (unmodified reference Evaluator / AverageMeter): tests/golden/,
scripts/make_golden_data.py: the reference (ga1i13o/DiffewS) ships no tests, golden vectors or fixtures for this path, and it
cannot be imported here (it needs diffusers==0.25.0, xformers, accelerate, matplotlib, detectron2 — none installed, no
network).  The arithmetic lives in the third-party dependency diffusers==0.25.0 (requirements.txt:2), restated from
its published architecture; the restatement is anchored on (i) the reference's own call sites cited per function,
(ii) exact agreement of the parameter counts and state-dict keys with SD-2.1 (UNet 865 910 724 + 23 360 conv_in_ref,
VAE 83 653 863), (iii) structural self-checks (k-shot fold == shot-major concat; scheduler step == negation; rthres
fp32 tie semantics; histc drops 255), all in tests/test_oracle.py.
"""
