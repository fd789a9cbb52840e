"""Asynchronous host->device staging of one training batch ahead of the step that consumes it.

Restates ``load_clips()`` of the reference loop (``app/avjepa/train.py:397-427``: ``.to(device,
non_blocking=True)`` of the clips and masks, and the implicit move of ``asgram`` that the reference leaves
to ``DataParallel.scatter``) with the copies issued on a dedicated CUDA stream, so the 234 MB of a
ViT-L batch cross PCIe while the previous step is still computing.  Host tensors should be pinned
(``DataLoader(pin_memory=True)`` in the reference, ``src/datasets/audiovideo_dataset.py:83``).
"""
import torch


def _to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(o, device) for o in obj)
    return obj


def _record(obj, stream):
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _record(o, stream)


class DevicePrefetcher(object):
    """Iterates ``batches`` (any iterable of nested tensors/lists/tuples) one element ahead on a copy
    stream.  Each yielded batch is safe to use on the current stream."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('DevicePrefetcher needs a CUDA device; avjepa_b200 has no CPU path')
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.next = None
        self._stage()

    def _stage(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.next = None
            return
        with torch.cuda.stream(self.copy_stream):
            self.next = _to_device(host, self.device)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.copy_stream)
        batch = self.next
        _record(batch, cur)
        self._stage()
        return batch
