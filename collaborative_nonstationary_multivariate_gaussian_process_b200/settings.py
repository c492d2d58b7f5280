"""Parity-critical constants of the SIM_code line, under the names the reference's modules import
(code/SIM_code/Utility/settings.py)."""
import torch

SELF_COVARIANCE_JITTER = 1e-6   # added to the diagonal of every self-covariance (kernels.py:35, 64)
VARIANCE_FLOOR = 1e-6           # replaces non-positive predictive variances (prediction.py) and scales logpdf1's jitter

# reference spellings
jitter, precision = SELF_COVARIANCE_JITTER, VARIANCE_FLOOR
torchType = torch.DoubleTensor  # the whole line computes in float64
