"""code/SIM_code/Utility/settings.py:1-6 (parity-critical constants)."""
import torch

jitter = 1e-6
torchType = torch.DoubleTensor
precision = 1e-6
