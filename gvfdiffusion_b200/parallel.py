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


class PipelinedExchange:
    """The data-parallel loop of SURVEY.md section 8(e) with its two exchanges overlapped with compute:
    rank `root` owns the objects (pinned host memory) and collects the frames; every step it uploads the
    NEXT step's conditioning of all ranks and sends each rank its share, while the frames of the PREVIOUS
    step travel back and are read to the host -- one batched NCCL point-to-point group per step on a side
    stream, double-buffered inputs and outputs, CUDA events between the side stream and the compute stream
    (no host synchronisation inside the loop).  With gloo / CPU tensors the same calls run synchronously
    (tests/test_parallel_cpu.py).

        ex = PipelinedExchange(in_specs, out_shape, device)
        ex.prime(host_inputs_of_step_0)            # list over ranks of dicts (root) / None
        for k in range(K):
            ex.post(k, host_inputs_of_step_k_plus_1)
            inp, out = ex.inputs(k), ex.output(k)  # device tensors of this rank for step k
            ... compute into `out` ...
            ex.done(k)
        ex.flush(K)                                 # frames of the last step have reached the root's host memory
        frames = ex.host_results                    # root: [world] pinned tensors (latest gathered step)
    """

    def __init__(self, in_specs, out_shape, device, out_dtype=torch.float32, root=0):
        self.rank, self.world, self.root = dist.get_rank(), dist.get_world_size(), root
        self.dev = torch.device(device)
        self.cuda = self.dev.type == "cuda"
        mk = lambda shape, dtype: torch.empty(shape, dtype=dtype, device=self.dev)
        self.names = sorted(in_specs)
        self.inp = [{n: mk(*in_specs[n]) for n in self.names} for _ in range(2)]
        self.out = [mk(out_shape, out_dtype) for _ in range(2)]
        self.is_root = self.rank == root
        if self.is_root:
            peers = [r for r in range(self.world) if r != root]
            self.stage = {r: {n: mk(*in_specs[n]) for n in self.names} for r in peers}
            self.gat = {r: mk(out_shape, out_dtype) for r in peers}
            pin = (lambda t: t.pin_memory()) if self.cuda else (lambda t: t)
            self.host_results = [pin(torch.empty(out_shape, dtype=out_dtype)) for _ in range(self.world)]
        if self.cuda:
            self.comm = torch.cuda.Stream(device=self.dev)
            self.xfer_done = [torch.cuda.Event(), torch.cuda.Event()]     # inputs of slot arrived, output of slot read out
            self.computed = [torch.cuda.Event(), torch.cuda.Event()]
        self.bytes_h2d = self.bytes_d2h = self.bytes_p2p = 0
        self._gather_pending = None

    # -- one exchange on the side stream: scatter into `slot_in` (if host_inputs / expected) and gather out of `slot_out`
    def _exchange(self, slot_in, host_inputs, slot_out, after=None):
        def body():
            ops = []
            if slot_in is not None:
                if self.is_root:
                    for r in range(self.world):
                        dst = self.inp[slot_in] if r == self.root else self.stage[r]
                        for n in self.names:
                            dst[n].copy_(host_inputs[r][n], non_blocking=True)
                            self.bytes_h2d += dst[n].numel() * dst[n].element_size()
                        if r != self.root:
                            for n in self.names:
                                ops.append(dist.P2POp(dist.isend, dst[n], r))
                                self.bytes_p2p += dst[n].numel() * dst[n].element_size()
                else:
                    for n in self.names:
                        ops.append(dist.P2POp(dist.irecv, self.inp[slot_in][n], self.root))
            if slot_out is not None:
                if self.is_root:
                    for r in self.gat:
                        ops.append(dist.P2POp(dist.irecv, self.gat[r], r))
                        self.bytes_p2p += self.gat[r].numel() * self.gat[r].element_size()
                else:
                    ops.append(dist.P2POp(dist.isend, self.out[slot_out], self.root))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                       # NCCL: the side stream waits, the host does not
            if slot_out is not None and self.is_root:
                for r in range(self.world):
                    src = self.out[slot_out] if r == self.root else self.gat[r]
                    self.host_results[r].copy_(src, non_blocking=True)
                    self.bytes_d2h += src.numel() * src.element_size()

        if not self.cuda:
            return body()
        if after is not None:
            self.comm.wait_event(after)
        with torch.cuda.stream(self.comm):
            body()

    def prime(self, host_inputs):
        """Inputs of step 0 (not part of a timed loop)."""
        self._exchange(0, host_inputs, None)
        if self.cuda:
            self.xfer_done[0].record(self.comm)
        self._gather_pending = None

    def post(self, k, next_host_inputs, scatter=True):
        """Start the transfers that overlap step k: conditioning of step k+1 -> slot (k+1)&1 (`scatter=False` on
        EVERY rank when there is no step k+1), frames of step k-1 <- the same slot.  Both wait for compute(k-1),
        the last user of that slot."""
        s = (k + 1) & 1
        gather = s if self._gather_pending == k - 1 else None
        self._exchange(s if scatter else None, next_host_inputs, gather,
                       after=self.computed[s] if (self.cuda and k > 0) else None)
        if self.cuda:
            self.xfer_done[s].record(self.comm)
        self._gather_pending = None

    def inputs(self, k):
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.xfer_done[k & 1])
        return self.inp[k & 1]

    def output(self, k):
        return self.out[k & 1]

    def done(self, k):
        if self.cuda:
            self.computed[k & 1].record(torch.cuda.current_stream())
        self._gather_pending = k

    def flush(self, k_next):
        """Gather the frames of step k_next - 1 (nothing else) and make the CURRENT stream wait for them."""
        if self._gather_pending == k_next - 1:
            s = (k_next - 1) & 1
            self._exchange(None, None, s, after=self.computed[s] if self.cuda else None)
            self._gather_pending = None
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm)
