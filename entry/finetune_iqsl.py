#!/usr/bin/env python
"""Adapter finetuning with the intensity-quantised structural loss — flags of the reference's finetune_iqsl.py:23-110.
Same loop as entry/finetune.py (frozen base under no_grad + OutputAdapter, device-side patch cropper, L1 + lambda_grad *
gradient loss) plus lambda_iqsl * iqsl_loss(pred, clean, t1, t2) (finetune_iqsl.py:291-383, :469-483; thresholds = the
(iqsl_q1, iqsl_q2) quantiles of the clean images, :258-288) and adapter-only checkpoints `epoch_adapter_only_XXX.pth`
(:114-132)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from entry import finetune  # noqa: E402

parser = finetune.parser
parser.set_defaults(arch='UNetImproved', log_name='UNetImproved_adapter_ft')       # finetune_iqsl.py:36-48
parser.add_argument('--lambda_iqsl', type=float, default=0.1)
parser.add_argument('--iqsl_q1', type=float, default=0.2)
parser.add_argument('--iqsl_q2', type=float, default=0.8)
parser.add_argument('--iqsl_tau', type=float, default=0.1)
parser.add_argument('--iqsl_margin', type=float, default=0.0)
parser.add_argument('--iqsl_max_images', type=int, default=50)
parser.add_argument('--iqsl_ce_factor', type=float, default=0.5)


if __name__ == "__main__":
    args, _ = parser.parse_known_args()
    finetune.main(args, iqsl=True)
