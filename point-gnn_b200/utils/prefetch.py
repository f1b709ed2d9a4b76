"""Graph generation of the NEXT batch overlapped with the model's forward pass of the current one.

The reference hides graph generation behind the network the same way: ``DataProvider`` (train.py:419-483)
hands ``fetch_data`` - which ends in ``graph_generate_fn`` (train.py:88-90) - to a pool of worker processes
(``Pool.apply_async``, train.py:430-455) and the training loop picks up finished batches.  Here both halves run
on the GPU, so the "workers" are a second CUDA stream: the host->device copy of batch i+1 and its graph build
(small, latency-bound kernels that leave most SMs idle) are issued on the side stream while the persistent
tensor-core kernels of batch i own the compute stream.  The single host round trip of the graph build (its
sizes) then blocks the host only on the side stream - the model keeps running.

    pf = GraphPrefetcher(graph_fn, graph_kwargs)
    ticket = pf.submit(points_host, intensity_host, frame_ptr_host)       # batch 0
    for i in range(n):
        intensity, coords, keypoints, edges = pf.collect(ticket)          # compute stream now waits for the graph
        logits, boxes = model.predict(intensity, coords, keypoints, edges, is_training=True)
        ...enqueue the device->host copies of the results...
        if i + 1 < n:
            ticket = pf.submit(*batch[i + 1])                             # overlaps with the predict above
        ...synchronise the compute stream, use the results...

Memory: the tensors of a ticket are allocated on the side stream and read on the compute stream; ``collect``
records that use with the caching allocator (``Tensor.record_stream``), so dropping them early is safe.
"""
import torch


class GraphTicket(object):
    __slots__ = ('intensity', 'coords', 'keypoints', 'edges', 'frame_ptrs', 'event', 'inputs')


class GraphPrefetcher(object):
    def __init__(self, graph_fn, graph_kwargs, device=None, stream=None):
        if not torch.cuda.is_available():
            raise RuntimeError('point-gnn_b200 needs a CUDA device (no CPU fallback)')
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.graph_fn = graph_fn
        self.graph_kwargs = dict(graph_kwargs)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)

    def submit(self, points_xyz, intensity, frame_ptr=None):
        """Start the copy (pinned host tensors / NumPy arrays are copied, CUDA tensors are used as they are) and
        the graph build of one batch on the side stream.  Returns when the build has been issued AND its sizes are
        known (the build's one host round trip) - work already queued on other streams keeps running meanwhile."""
        t = GraphTicket()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))      # inputs produced on the caller's stream
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            xyz = self._to_device(points_xyz, torch.float32)
            t.intensity = self._to_device(intensity, torch.float32)
            fp = None if frame_ptr is None else self._to_device(frame_ptr, torch.int32)
            t.inputs = (xyz, fp)
            out = self.graph_fn(xyz, frame_ptr=fp, return_frame_ptr=True, **self.graph_kwargs)
            t.coords, t.keypoints, t.edges, t.frame_ptrs = out
            t.event = torch.cuda.Event()
            t.event.record(self.stream)
        return t

    def collect(self, ticket):
        """Make the current stream wait for the ticket's graph; -> (intensity, coords, keypoints, edges)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ticket.event)
        for x in [ticket.intensity] + list(ticket.coords) + list(ticket.keypoints) + list(ticket.edges):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                x.record_stream(cur)
        return ticket.intensity, ticket.coords, ticket.keypoints, ticket.edges

    def _to_device(self, x, dtype):
        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                return x.to(dtype=dtype).contiguous()
            return x.to(self.device, dtype=dtype, non_blocking=True)
        import numpy as np
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device, dtype=dtype, non_blocking=True)
