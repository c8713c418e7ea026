"""Host-side helpers mirrored from the reference's utils/utils.py that sit on the hot path's callers."""
import torch
import torch.nn as nn


class MultiCropWrapper(nn.Module):
    """utils/utils.py:564-609: one backbone forward per group of same-resolution views (or the fixed
    [1,2,4,6,8,10] grouping of the multi-frame-rate recipe when vary_fr), concatenate, then the head."""

    def __init__(self, backbone, head, vary_fr=False):
        super().__init__()
        if hasattr(backbone, 'fc'):
            backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head
        self.vary_fr = vary_fr

    def forward(self, x, **kwargs):
        if not isinstance(x, list):
            x = [x]
        idx_crops = torch.cumsum(torch.unique_consecutive(
            torch.tensor([inp.shape[-1] for inp in x]), return_counts=True)[1], 0)
        if self.vary_fr:
            if len(x) == 10:
                idx_crops = [1, 2, 4, 6, 8, 10]
            elif len(x) == 2:
                idx_crops = [1, 2]
        start_idx, output = 0, None
        for end_idx in idx_crops:
            _out = self.backbone(torch.cat(x[start_idx: end_idx]), **kwargs)
            output = _out if output is None else torch.cat((output, _out))
            start_idx = end_idx
        return self.head(output)
