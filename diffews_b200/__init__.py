"""diffews_b200 — B200-native (sm_100a) hot path of DiffewS behind the reference's own Python API.

Sub-modules:
  _lib / ops              C-ABI binding (libdiffews_b200.so) and torch-tensor front end
  unet / vae              weight preparation + layer schedules that drive the kernels
  attention_processor     MyAttention-style KV-bank processor (reference: diffews/models/attention_processor.py)
  pipeline                MarigoldPipelineRGBLatentNoise drop-in (reference: diffews/marigold_pipeline_rgb_latent_noise.py)
  evaluation              Evaluator / AverageMeter drop-ins with integer counts
"""
__version__ = "0.1.0"
