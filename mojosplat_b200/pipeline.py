"""Software-pipelined multi-view rendering on one GPU (additive; BASELINE.json configs 3-4).

A frame has two halves with opposite characters: ``begin`` (projection, depth sort of the N
Gaussians, count + scan: ~10 small latency-bound kernels that leave most of the chip idle) and
``end`` (emission, tile sort, tile ranges, rasterization: the rasterizer fills every SM).  Between
them sits the frame's only host read-back (M).  ``FramePipeline`` keeps ``depth`` frames in flight on
``depth`` CUDA streams with one workspace each, so that begin(k+1) runs on the GPU while end(k) is
executing and the host wait for M(k+1) is hidden behind end(k).  Results are identical to
frame-by-frame rendering (same kernels, same order inside a frame).
"""
from __future__ import annotations

from ctypes import byref, c_size_t
from typing import Sequence

import torch

from . import _lib
from .rasterization import RASTER_MODES
from .utils import Camera


class FramePipeline:
    def __init__(self, device, N: int, W: int, H: int, channels: int = 3, tile_size: int = 16,
                 semantics: int = _lib.SEM_TORCH, depth: int = 2, m_capacity: int | None = None,
                 raster_mode: str = "fast"):
        self.dev = torch.device(device)
        self.L = _lib.require_device(self.dev)
        self.N, self.W, self.H, self.C, self.ts = int(N), int(W), int(H), int(channels), int(tile_size)
        self.semantics, self.flags = semantics, RASTER_MODES[raster_mode]
        self.depth = depth
        self.m_cap = int(m_capacity) if m_capacity else 8 * self.N + 4096
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(self.dev) for _ in range(depth)]
            self.events = [torch.cuda.Event() for _ in range(depth)]
            self.info = [torch.zeros(32, dtype=torch.uint8).pin_memory() for _ in range(depth)]
            self.ws = [self._alloc(self.m_cap) for _ in range(depth)]
        self.last_M = 0

    def _alloc(self, m_cap: int) -> torch.Tensor:
        nbytes = self.L.bsplat_render_workspace_bytes(self.N, m_cap, self.W, self.H, self.ts)
        return torch.empty(nbytes, dtype=torch.uint8, device=self.dev)

    # -- the two halves ---------------------------------------------------------------------
    def _begin(self, slot: int, g, cam_struct) -> None:
        s = self.streams[slot]
        rc = self.L.bsplat_render_begin(self.N, _lib.ptr(g[0]), _lib.ptr(g[1]), _lib.ptr(g[2]), _lib.ptr(g[3]),
                                        byref(cam_struct), self.ts, self.semantics, _lib.ptr(self.ws[slot]),
                                        self.ws[slot].numel(), self.info[slot].data_ptr(), s.cuda_stream)
        _lib.check(rc, "bsplat_render_begin")
        self.events[slot].record(s)

    def _end(self, slot: int, g, cam_struct, background, image) -> int:
        self.events[slot].synchronize()  # host waits for begin(k) only; the other stream keeps the GPU busy
        M = int(_lib.BsplatBinInfo.from_buffer_copy(self.info[slot].numpy().tobytes()).n_isect)
        s = self.streams[slot]
        needed = c_size_t(0)
        for attempt in range(2):
            rc = self.L.bsplat_render_end(self.N, M, _lib.ptr(g[4]), _lib.ptr(g[3]), self.C, byref(cam_struct),
                                          _lib.ptr(background), self.ts, self.semantics, self.flags,
                                          _lib.ptr(image), _lib.ptr(self.ws[slot]), self.ws[slot].numel(),
                                          byref(needed), s.cuda_stream)
            if rc == _lib.E_WORKSPACE and attempt == 0:
                # rare: this view produced more intersections than the workspace was sized for.
                # Grow the slot and redo its begin half (the N-part lives in the old buffer).
                s.synchronize()
                self.m_cap = int(M * 1.3) + 4096
                self.ws[slot] = self._alloc(self.m_cap)
                self._begin(slot, g, cam_struct)
                self.events[slot].synchronize()
                continue
            break
        _lib.check(rc, "bsplat_render_end")
        self.last_M = M
        return M

    # -- driver -----------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, means3d, scales, quats, opacities, features, cameras: Sequence[Camera], background,
               out: torch.Tensor | None = None, scene_of=None) -> torch.Tensor:
        """Render all cameras; returns images [n, H, W, C] (valid on the current stream on return).
        If ``out`` has fewer than n slots it is used as a ring (frame k -> out[k % len(out)]).
        ``scene_of`` (optional): callable k -> (means3d, scales, quats, opacities, features) giving frame k
        its own Gaussian set of the same N (dynamic scenes; bench.py rotates copies so that the inputs
        touched between two uses of one copy exceed the L2)."""
        def prep(t):
            q = [_lib.as_f32(t[0], "means3d"), _lib.as_f32(t[1], "scales"), _lib.as_f32(t[2], "quats"),
                 _lib.as_f32(t[3], "opacities").reshape(-1), _lib.as_f32(t[4], "features")]
            assert q[0].shape[0] == self.N and q[4].shape == (self.N, self.C)
            return q
        g0 = prep((means3d, scales, quats, opacities, features))
        gs = [g0] * len(cameras) if scene_of is None else [prep(scene_of(k)) for k in range(len(cameras))]
        bg = _lib.as_f32(background, "background").to(self.dev)
        n = len(cameras)
        if out is None:
            out = torch.empty((n, self.H, self.W, self.C), dtype=torch.float32, device=self.dev)
        cams = [_lib.camera_struct(c) for c in cameras]
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.device(self.dev):
            for s in self.streams:
                s.wait_stream(cur)  # inputs were produced on the caller's stream
            if n > 0:
                self._begin(0, gs[0], cams[0])
            for k in range(n):
                slot = k % self.depth
                self._end(slot, gs[k], cams[k], bg, out[k % out.shape[0]])
                # enqueue the next frame's first half right behind: it overlaps end(k) on the GPU.
                # With depth 2 the slot of frame k+1 finished end(k-1) long ago in stream order.
                if k + 1 < n:
                    self._begin((k + 1) % self.depth, gs[k + 1], cams[k + 1])
            for s in self.streams:
                cur.wait_stream(s)
        return out
