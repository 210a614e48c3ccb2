#!/usr/bin/env python
"""Counterpart of the reference's project/run_machine_translation.py (same flags: --use-flash-attention,
--model-max-length), device-resident and data-parallel:

    python project/run_machine_translation.py --use-flash-attention True --model-max-length 512
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        project/run_machine_translation.py --use-flash-attention True --model-max-length 512
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from fa_b200.mt.train import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
