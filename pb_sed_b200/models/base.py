"""``pb_sed.models.base.SoundEventModel`` surface (pb_sed/models/base/model.py:9-42) on a minimal
stand-in for ``padertorch.Model`` (padertorch is not importable here; when it is, ``Model`` can be
swapped for ``padertorch.Model`` without touching the subclasses)."""
import abc

import numpy as np
import torch
from torch import nn


class Model(nn.Module):
    """the slice of padertorch.Model the hot path uses: forward / review / example_to_device /
    modify_summary (pb_sed/models/base/inference.py:130; trainer step)."""

    def example_to_device(self, example, device=None):
        out = {}
        for key, value in example.items():
            if isinstance(value, np.ndarray) and value.dtype.kind in 'fiub':
                value = torch.from_numpy(value)
            if torch.is_tensor(value):
                value = value.to(device, non_blocking=True)
            out[key] = value
        return out

    def review(self, inputs, outputs):
        raise NotImplementedError

    def modify_summary(self, summary):
        return summary


class SoundEventModel(Model, abc.ABC):
    def __init__(self, *, labelwise_metrics=(), label_mapping=None, test_labels=None):
        super().__init__()
        self.labelwise_metrics = labelwise_metrics
        self.label_mapping = label_mapping
        self.test_labels = test_labels

    @abc.abstractmethod
    def tagging(self, inputs, **params):
        pass

    @abc.abstractmethod
    def boundaries_detection(self, inputs, **params):
        pass

    @abc.abstractmethod
    def sound_event_detection(self, inputs, **params):
        pass

    def modify_summary(self, summary):
        """mean of the per-batch scalars (model.py:28-31); image grids and the buffered-score
        metrics (model.py:33-88) are host-side validation code outside the hot path."""
        for key, scalar in summary.get('scalars', {}).items():
            summary['scalars'][key] = np.mean(scalar)
        return summary
