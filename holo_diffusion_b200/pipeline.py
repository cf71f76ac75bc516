"""Double-buffered end-to-end view rendering: inputs in (pinned) HOST memory, images back in host memory.

``HoloDiffusionModel.forward`` ends with the reference's range asserts (holo_diffusion_model.py:381,426,428), i.e. a
16-byte device-to-host read that blocks the host until the view is finished -- so a caller that uploads the next
33.5 MB grid only after ``forward`` returns serialises copy and compute (0.8 ms of a 10.7 ms step on one B200, more
when 8 ranks share the host's memory system).  ``ViewStream`` issues the upload of view i+1 on a copy stream BEFORE it
runs view i, into the other half of a device double buffer; the asserts still fire, per view, in order.

    vs = ViewStream(model)
    for preds, image_host in vs.render(zip(host_grids, host_cameras)):
        ...

or, step by step (bench.py): ``t = vs.prefetch(grid_host, cam_host)`` ... ``preds = vs.run(t)``.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Tuple

import torch

from .cameras import PerspectiveCameras


class _Ticket:
    __slots__ = ("slot", "ready")

    def __init__(self, slot: int, ready: torch.cuda.Event):
        self.slot, self.ready = slot, ready


class ViewStream:
    def __init__(self, model, device=None, n_slots: int = 2):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        C, R = model.feature_size, model.resol
        self.n_slots = n_slots
        self._grid = [torch.empty(1, C, R, R, R, device=self.device) for _ in range(n_slots)]
        self._cam = [PerspectiveCameras(torch.ones(1, 2), torch.zeros(1, 2), torch.eye(3)[None], torch.zeros(1, 3)).to(self.device)
                     for _ in range(n_slots)]
        self._free = [None] * n_slots          # event: the compute stream finished reading slot s
        self._copy = torch.cuda.Stream(device=self.device)
        self._next = 0

    def prefetch(self, grid_host: torch.Tensor, cam_host: PerspectiveCameras) -> _Ticket:
        """Queue the host-to-device copies of one view's inputs on the copy stream (asynchronous when the host
        tensors are pinned)."""
        s = self._next
        self._next = (s + 1) % self.n_slots
        with torch.cuda.stream(self._copy):
            if self._free[s] is not None:
                self._copy.wait_event(self._free[s])
            self._grid[s].copy_(grid_host, non_blocking=True)
            c = self._cam[s]
            c.R.copy_(cam_host.R[:1], non_blocking=True)
            c.T.copy_(cam_host.T[:1], non_blocking=True)
            c.focal_length.copy_(cam_host.focal_length[:1], non_blocking=True)
            c.principal_point.copy_(cam_host.principal_point[:1], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy)
        return _Ticket(s, ready)

    def run(self, ticket: _Ticket, **forward_kwargs):
        """``model.forward`` on a prefetched view (blocks the host on the view's range asserts, like forward)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ticket.ready)
        preds = self.model(camera=self._cam[ticket.slot], voxel_features=self._grid[ticket.slot], **forward_kwargs)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._free[ticket.slot] = ev
        return preds

    def render(self, views: Iterable[Tuple[torch.Tensor, PerspectiveCameras]],
               image_host: Optional[torch.Tensor] = None) -> Iterator[Tuple[dict, Optional[torch.Tensor]]]:
        """Render a sequence of (host grid, host camera) pairs; the upload of view i+1 overlaps the compute of view i.
        image_host: optional pinned (5, H, W) buffer that receives [rgb | depth | mask] of every view."""
        it = iter(views)
        try:
            nxt = self.prefetch(*next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur = nxt
            try:
                nxt = self.prefetch(*next(it))
            except StopIteration:
                nxt = None
            preds = self.run(cur)
            if image_host is not None:
                img = torch.cat([preds["images_render"][0], preds["depths_render"][0], preds["masks_render"][0]], 0)
                image_host.copy_(img, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
            yield preds, image_host
