"""Device selection with the reference's contract (finenvs/device_utils.py:4-9).

`set_device` keeps the reference behaviour (agents call it and may run on CPU); the env itself has no
CPU path and uses `require_cuda_device`, which raises instead of falling back.
"""
import torch


def set_device(device_id: int) -> str:
    use_cuda = device_id >= 0 and torch.cuda.is_available()
    if not use_cuda:  # the message is part of the reference's observable behaviour (device_utils.py:8)
        print("WARNING: PyTorch not recognizing CUDA device -> forcing CPU...")
    return f"cuda:{device_id}" if use_cuda else "cpu"


def require_cuda_device(device_id: int) -> str:
    device = set_device(device_id)
    if device == "cpu":
        raise RuntimeError(
            "finenvs_b200.TimeSeriesEnv runs only on a CUDA device (sm_100a kernel, no CPU fallback); "
            f"device_id={device_id}, torch.cuda.is_available()={torch.cuda.is_available()}"
        )
    return device
