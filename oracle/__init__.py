"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch fp32 / numpy) of the reference's algorithm for the DiffewS hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only as the
checker.  The product package `diffews_b200` never imports `oracle`.

PARITY — pinned against outputs of the reference itself, run in the build container: scripts/make_golden_*.py execute the
reference's own files UNMODIFIED (only the names they import from packages that are not installed — diffusers, xformers,
accelerate, matplotlib, detectron2 ... — are replaced by stand-ins that carry no arithmetic of the path) and commit what
they return under tests/golden/: the whole `test_diffusion` loop (eval_loop_reference.json), the pipeline's __call__ /
single_infer (pipeline_reference.json), MyUNet2DConditionModel.forward on a hand-assembled instance
(unet_wiring_reference.json), the KV-bank attention processors (attn_reference.json), the scheduler tables
(scheduler_reference.json), Evaluator / AverageMeter (metric_reference.json) and the datasets (data_layer.json).
tests/test_oracle.py and tests/test_data_layer.py hold this oracle to them.  NOT backed by a reference run: the insides of the
diffusers modules the reference instantiates (UNet blocks, VAE encoder / decoder, DDIMScheduler.step) — diffusers==0.25.0
(requirements.txt:2) is absent, so oracle/sd21.py restates its published architecture, anchored on (i) the reference's call
sites cited per function, (ii) exact agreement of the parameter counts and state-dict keys with SD-2.1 (UNet 865 910 724 +
23 360 conv_in_ref, VAE 83 653 863), (iii) structural self-checks in tests/test_oracle.py.
"""
