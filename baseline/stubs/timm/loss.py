"""`timm.loss` classes main.py:29 imports (classification path; the multi-task path uses mtl_loss_schemes)."""
import torch
import torch.nn.functional as F


class LabelSmoothingCrossEntropy(torch.nn.Module):
    def __init__(self, smoothing=0.1):
        super().__init__()
        self.smoothing, self.confidence = smoothing, 1.0 - smoothing

    def forward(self, x, target):
        logprobs = F.log_softmax(x, dim=-1)
        nll = -logprobs.gather(dim=-1, index=target.unsqueeze(1)).squeeze(1)
        return (self.confidence * nll + self.smoothing * -logprobs.mean(dim=-1)).mean()


class SoftTargetCrossEntropy(torch.nn.Module):
    def forward(self, x, target):
        return torch.sum(-target * F.log_softmax(x, dim=-1), dim=-1).mean()
