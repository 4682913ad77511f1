"""Layout conventions of the reference: tensors are [B,C,H,W] fp32; arrays given as HWC
(last dim 1 or 3) are moved to CHW and a batch axis is added; a `dp.tensor` carries a tag that
suppresses re-batching (dprox/utils/misc.py:42-96, dprox/utils/containar.py:40-48)."""
import numpy as np
import torch


def is_dp_tensor(x) -> bool:
    return getattr(x, "is_dp_tensor", False) is True


def tensor(*args, **kwargs) -> torch.Tensor:
    out = torch.tensor(*args, **kwargs)
    out.is_dp_tensor = True
    return out


def array(*args, **kwargs) -> np.ndarray:
    return np.array(*args, **kwargs)


def to_torch_tensor(x, batch: bool = False) -> torch.Tensor:
    if is_dp_tensor(x):
        return x
    if isinstance(x, torch.Tensor):
        out = x
    elif isinstance(x, np.ndarray):
        out = torch.tensor(x.copy())
    else:
        out = torch.tensor(x)
    if batch:
        if out.ndim == 3 and out.shape[2] in (1, 3):
            out = out.permute(2, 0, 1)
        if out.ndim < 4:
            out = out.unsqueeze(0)
    out.is_dp_tensor = True
    return out


def to_ndarray(x, debatch: bool = False, squeeze: bool = False) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        out = x.detach().cpu().numpy()
    elif isinstance(x, np.ndarray):
        out = x.astype("float32")
    else:
        out = np.array(x)
    if debatch:
        if out.ndim == 4:
            out = out.squeeze(0)
        if out.ndim == 3:
            if out.shape[0] in (1, 3):
                out = out.transpose(1, 2, 0)
            if out.shape[2] == 1 and squeeze:
                out = out.squeeze(2)
    return out


def as_bchw(t: torch.Tensor) -> torch.Tensor:
    """View a <=4-D tensor as [B,C,H,W] (the kernels are 4-D): [B,H,W] -> [B,1,H,W], [B,W] -> [B,1,1,W]
    (a batched tensor keeps its leading batch axis, like the reference's fft over dims [-2,-1])."""
    if t.ndim == 3:
        t = t.unsqueeze(1)
    elif t.ndim == 2:
        t = t.unsqueeze(1).unsqueeze(1)
    elif t.ndim == 1:
        t = t.reshape(1, 1, 1, -1)
    elif t.ndim == 0:
        t = t.reshape(1, 1, 1, 1)
    if t.ndim != 4:
        raise ValueError(f"expected a tensor with at most 4 dims, got shape {tuple(t.shape)}")
    return t
