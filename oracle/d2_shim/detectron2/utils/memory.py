"""detectron2.utils.memory.retry_if_cuda_oom: on CPU tensors the wrapper is a plain call."""
from functools import wraps


def retry_if_cuda_oom(func):
    @wraps(func)
    def wrapped(*args, **kwargs):
        return func(*args, **kwargs)
    return wrapped
