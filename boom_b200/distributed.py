"""Row sharding of one model over the GPUs of a box: one process per GPU (torch.distributed, NCCL),
contiguous row blocks, ONE all-reduce of the packed statistics per Gibbs iteration.

The reference parallelises the same loop over host threads: assign_data_to_workers gives worker w the
contiguous block [w * floor(n / W), (w + 1) * floor(n / W)) and the last worker the remainder
(Models/PosteriorSamplers/Imputer.hpp:348-375), and the per-worker statistics are summed by combine()
(BinomialLogitAuxmixSampler.cpp:44-49, WeightedRegressionModel.cpp:79-87).  Here a worker is a GPU and
combine() is ncclAllReduce(sum) over NVLink on [p*p | p | 4] doubles (boomgpu_suf_len).  The Philox
counters are keyed by the GLOBAL row index (boomgpu_set_row_offset), so the latent draws do not depend
on the number of shards; only the floating-point summation order does.

Every rank then runs the same host small-state step on the same all-reduced statistics with the same
sampler seed, so beta and gamma stay identical on all ranks without a broadcast."""


def shard_range(n, world, rank):
    """Rows [row0, row1) of rank `rank` out of `world` (Imputer.hpp:348-375: equal blocks, remainder to the last)."""
    if world < 1 or not 0 <= rank < world or n < 0:
        raise ValueError("bad shard request n=%r world=%r rank=%r" % (n, world, rank))
    rows = n // world
    row0 = rank * rows
    return row0, (n if rank == world - 1 else row0 + rows)


class _DevicePointer:
    """Zero-copy torch view of a device pointer (__cuda_array_interface__ v2)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class StatisticsAllReduce:
    """The hook model.set_allreduce() takes: sums count doubles at a device pointer over the process group,
    in place, ordered after the device step on `stream` (the stream given to model.set_stream)."""

    def __init__(self, stream, device, group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.stream, self.device, self.group = stream, device, group
        self._views = {}
        self.calls = 0
        self.bytes = 0

    def __call__(self, ptr, count):
        torch = self._torch
        view = self._views.get((ptr, count))
        if view is None:
            view = self._views[(ptr, count)] = torch.as_tensor(_DevicePointer(ptr, count), device=self.device)
        with torch.cuda.stream(self.stream):
            self._dist.all_reduce(view, group=self.group)
        self.calls += 1
        self.bytes += 8 * int(count)


def attach(model, n_total, stream, device, rank=None, world=None, group=None, native=False):
    """Binds `model` (holding this rank's rows of an n_total-row data set) to its shard: device, stream,
    global row offset and the all-reduce.  native=False: a torch.distributed hook on the device buffer;
    native=True: the library joins its own NCCL communicator (boomgpu_comm_init; the 128-byte id is made on rank 0 and
    broadcast through torch.distributed) and all-reduces inside the C ABI step.  Returns (row0, row1, hook or None)."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    row0, row1 = shard_range(n_total, world, rank)
    model.set_device(device.index if device.index is not None else 0)
    model.set_row_offset(row0)
    model.set_stream(stream.cuda_stream)
    hook = None
    if world > 1 and native:
        box = [type(model).comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        model.set_communicator(box[0], world, rank)
    elif world > 1:
        hook = StatisticsAllReduce(stream, device, group)
        model.set_allreduce(hook)
    return row0, row1, hook
