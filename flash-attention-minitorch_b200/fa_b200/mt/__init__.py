"""Decoder-only LM training on device-resident tensors with the B200 flash-attention kernels
(BASELINE.json configs[4]; reference project/run_machine_translation.py).

  ops.py    device-resident operators (attention = this repo's kernels)
  model.py  DecoderLM / TransformerLayer / MultiHeadAttention / FeedForward / LayerNorm1d / Linear / Embedding
  optim.py  Adam exactly as the reference writes it (minitorch/optim.py:30-76)
  train.py  training loop, synthetic token batches, data-parallel gradient all-reduce (NCCL over NVLink)
"""
