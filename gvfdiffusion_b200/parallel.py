"""Object-level data parallelism (one process per GPU, torch.distributed).

Objects are independent through sampler, decoder and renderer, so the path
shards with NO data-path collective: rank r takes objects r, r+W, r+2W, ...  The only exchanges are
the optional scatter of per-object conditioning from rank 0 and the gather of rendered frames --
both plain NCCL point-to-point / gather calls over NVLink (gloo on CPU for the tests).
The reference itself runs every object on every rank (inference_dpm_latent.py:159 with an
un-sharded loader, dataset_latent_inference.py:39-57)."""
import torch
import torch.distributed as dist


def object_shard(num_objects, rank, world):
    """Indices of the objects rank `rank` processes (round-robin)."""
    return list(range(rank, num_objects, world))


def owner(obj_index, world):
    return obj_index % world


def scatter_conditioning(tensors_per_object, num_objects, src=0, device=None):
    """Rank `src` holds a list (len num_objects) of dicts of tensors; every rank returns the dicts of
    its own shard {object_index: dict}.  Other ranks pass None.  Shapes/dtypes travel as objects first."""
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        meta[0] = [{k: (tuple(v.shape), v.dtype) for k, v in d.items()} for d in tensors_per_object]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for i in range(num_objects):
        dst = owner(i, world)
        if rank == src:
            if dst == src:
                out[i] = {k: (v.to(device) if device is not None else v) for k, v in tensors_per_object[i].items()}
            else:
                for k in sorted(tensors_per_object[i]):
                    t = tensors_per_object[i][k]
                    dist.send((t.to(device) if device is not None else t).contiguous(), dst=dst)
        elif rank == dst:
            d = {}
            for k in sorted(meta[0][i]):
                shape, dtype = meta[0][i][k]
                buf = torch.empty(shape, dtype=dtype, device=device)
                dist.recv(buf, src=src)
                d[k] = buf
            out[i] = d
    return out


def gather_results(results, num_objects, dst=0):
    """results: {object_index: tensor} of this rank's shard (same shape on every rank).  Returns on rank
    `dst` the list of all objects' tensors in object order, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    out = [None] * num_objects if rank == dst else None
    for i in range(num_objects):
        src = owner(i, world)
        if src == dst:
            if rank == dst:
                out[i] = results[i]
        elif rank == src:
            dist.send(results[i].contiguous(), dst=dst)
        elif rank == dst:
            ref = next(iter(results.values())) if results else None
            if ref is None:
                raise RuntimeError("gather_results needs at least one local result to size the buffers")
            buf = torch.empty_like(ref)
            dist.recv(buf, src=src)
            out[i] = buf
    return out
