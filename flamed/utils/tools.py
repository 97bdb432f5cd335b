"""The two helpers of the reference's flamed/utils/tools.py that are live on the inference path."""
import torch
import torch.nn.functional as F


def get_mask_from_lengths(lengths, max_len=None):
    """(B,) lengths -> (B, max_len) bool, True = padding (reference tools.py:91-99)."""
    if max_len is None:
        max_len = int(lengths.max().item())
    steps = torch.arange(max_len, device=lengths.device)
    return steps[None, :] >= lengths[:, None]


def pad(tensors, max_len=None):
    """zero-pad a list of (T,) or (T,H) tensors along dim 0 and stack (reference tools.py:299-317)."""
    if not max_len:
        max_len = max(t.size(0) for t in tensors)
    out = []
    for t in tensors:
        extra = max_len - t.size(0)
        out.append(F.pad(t, (0, extra) if t.dim() == 1 else (0, 0, 0, extra)))
    return torch.stack(out)
