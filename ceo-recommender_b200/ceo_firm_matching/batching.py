"""Device-resident batch pipeline.

The reference iterates a ``DataLoader`` over a map-style dataset whose ``__getitem__`` builds a 6-key dict
per sample (``data.py:190-198``); >90 % of its wall-clock at the shipped configs is that per-sample path plus
six small H2D copies per step (``training.py:42``).  Here the dataset's tensor dict is uploaded once and every
batch is an index-select on the device.  The batch *composition* is still decided by the loader's own
``batch_sampler`` (same sampler, same ``drop_last``); before iterating it one int64 is drawn from the loader's
generator (the global one when none is set), as ``DataLoader.__iter__`` does for its ``_base_seed``, so that after the
same ``torch.manual_seed`` the shuffles are the ones ``for batch in loader`` produces.
"""
from typing import Dict, Iterator, List, Sequence

import torch
from torch.utils.data import DataLoader


def _tensor_dict(loader: DataLoader, keys: Sequence[str]):
    data = getattr(loader.dataset, "data", None)
    if isinstance(data, dict) and all(isinstance(data.get(k), torch.Tensor) for k in keys):
        return data
    return None


def device_batches(loader: DataLoader, keys: Sequence[str], device: torch.device) -> Iterator[List[torch.Tensor]]:
    """Yield each batch of ``loader`` as a list of device tensors in ``keys`` order."""
    data = _tensor_dict(loader, keys)
    if data is None or loader.batch_sampler is None:
        for batch in loader:                                   # generic loaders: plain per-batch upload
            yield [batch[k].to(device, non_blocking=True) for k in keys]
        return
    cache: Dict[str, torch.Tensor] = getattr(loader, "_cfm_device_cache", None)
    stamp = tuple((data[k].data_ptr(), data[k]._version, tuple(data[k].shape)) for k in keys)   # source replaced / mutated?
    if cache is None or getattr(loader, "_cfm_device_stamp", None) != stamp or any(cache[k].device != device for k in keys):
        cache = {k: data[k].to(device) for k in keys}
        loader._cfm_device_cache, loader._cfm_device_stamp = cache, stamp
    # _BaseDataLoaderIter.__init__ draws its base seed from the loader's generator before the sampler draws its own
    torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
    for idx in loader.batch_sampler:
        sel = torch.as_tensor(idx, dtype=torch.long).to(device, non_blocking=True)
        yield [cache[k].index_select(0, sel) for k in keys]
