"""One training step of the EVE hot path as a replayable CUDA graph.

The step (time-batched ``EVE.forward``, ``full_loss.backward()``, gradient packing, optional
NCCL allreduce, fused clip + Adam) launches ~1.5k kernels from Python; captured once, a replay
costs one ``cudaGraphLaunch``.  Everything data-dependent stays outside the capture:
  * inputs are copied into static device buffers before each replay (H2D straight from the
    caller's pinned tensors);
  * the per-clip kappa augmentation draws (np.random, eve.py:468-469) are made on the host in
    the reference's order and copied into static [B, 2] buffers;
  * the Adam step count lives on the device (eve_adam_params.step_dev);
  * batches may carry their frames undecoded-to-float ('eyes_frames' / 'screen_frames' uint8 in the
    decoder's layout, eve_b200/input_pipeline.py): they cross PCIe as bytes and one kernel per
    stream converts them straight into the static float buffers;
  * ``prefetch()`` stages the NEXT batch host->device on a second stream while the current replay
    computes; the step then starts with a device-to-device copy into the static buffers.
The math is the eager path's, kernel for kernel -- tests/test_gpu_graph.py holds the two
bit-identical.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import input_pipeline as IP
from . import lib as L


class GraphedTrainStep(object):
    KAPPA_RING = 4

    def __init__(self, model, trainer, example_inputs, current_epoch=0.0, warmup=3,
                 tag='train', capture=True):
        self.model = model
        self.trainer = trainer
        self.tag = tag
        self.epoch = float(current_epoch)
        dev = trainer.device
        self.static_in = {k: torch.empty(v.shape, dtype=v.dtype, device=dev)
                          for k, v in example_inputs.items()}
        B = next(iter(example_inputs.values())).shape[0]
        # ring of pinned kappa buffers: the H2D copy of step i is asynchronous, so the host may not
        # rewrite a buffer before the copy that reads it has run (event per slot)
        self.kappa_host = [{s: torch.empty((B, 2), dtype=torch.float32).pin_memory()
                            for s in ('left', 'right')} for _ in range(self.KAPPA_RING)]
        self.kappa_done = [None] * self.KAPPA_RING
        self.kappa_slot = 0
        # recorded after the batch has been copied into the static buffers: callers that reuse
        # their pinned input tensors wait on it (inputs_consumed.synchronize()) before rewriting them
        self.inputs_consumed = torch.cuda.Event()
        model.kappa_buffers = {s: torch.zeros((B, 2), dtype=torch.float32, device=dev)
                               for s in ('left', 'right')}
        self.batch = B
        self.loss = None
        self.graph = None
        self._pinned = False
        # input prefetch: staging buffers filled on a copy stream (see prefetch())
        self.stage_in = None
        self.raw_dev = {}               # device copies of uint8 frame tensors handed over on the host
        self.copy_stream = None
        self.stage_ready = None
        self.stage_free = None
        self._load(example_inputs)
        # warm-up on a side stream (allocator pools, lazy kernel attributes, NCCL channels)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if capture:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._eager()
            torch.cuda.synchronize(dev)
            # capturing ran the host side of one step without executing it: the device-side
            # counter is the truth
            if trainer.step_dev is not None:
                trainer.steps = int(trainer.step_dev)
            # the graph's kernels hold pointers into the shared scratch buffers (lib.workspace):
            # from now on an outgrown buffer must stay allocated
            L.pin_workspaces()
            self._pinned = True

    def _eager(self):
        out = self.model({self.tag: dict(self.static_in)}, current_epoch=self.epoch)
        loss = out['full_loss']
        self.trainer.step(loss)
        return loss.detach()

    def _raw_on_device(self, k, v):
        if v.is_cuda:
            return v
        buf = self.raw_dev.get(k)
        if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
            buf = self.raw_dev[k] = torch.empty(v.shape, dtype=v.dtype, device=self.trainer.device)
        buf.copy_(v, non_blocking=True)
        return buf

    def _load(self, inputs):
        fpc = inputs.get('frames_per_clip')
        if fpc is not None:
            fpc = self._raw_on_device('frames_per_clip', fpc)
        for k, v in inputs.items():
            if k == 'eyes_frames':      # uint8 [B,T,H,2*ew,3] -> both static eye-patch buffers
                IP.preprocess_eye_frames(self._raw_on_device(k, v), fpc,
                                         out=(self.static_in['left_eye_patch'],
                                              self.static_in['right_eye_patch']))
            elif k == 'screen_frames':  # uint8 [B,T,72,128,3] -> the static screen_frame buffer
                IP.preprocess_screen_frames(self._raw_on_device(k, v), fpc,
                                            out=self.static_in['screen_frame'])
            elif k != 'frames_per_clip':
                self.static_in[k].copy_(v, non_blocking=True)
        self.inputs_consumed.record(torch.cuda.current_stream(self.trainer.device))
        left, right = self.model.draw_kappas(self.batch)
        slot = self.kappa_slot
        self.kappa_slot = (slot + 1) % self.KAPPA_RING
        if self.kappa_done[slot] is not None:
            self.kappa_done[slot].synchronize()     # the copy that last read this slot has run
        host = self.kappa_host[slot]
        host['left'].copy_(torch.from_numpy(left.astype(np.float32)))
        host['right'].copy_(torch.from_numpy(right.astype(np.float32)))
        for s in ('left', 'right'):
            self.model.kappa_buffers[s].copy_(host[s], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.trainer.device))
        self.kappa_done[slot] = ev

    def prefetch(self, inputs):
        """Start copying the NEXT step's inputs (pinned host tensors) into device staging buffers
        on a second stream; it overlaps whatever the main stream is computing.  The following
        ``__call__(None)`` consumes the staged batch."""
        dev = self.trainer.device
        if self.stage_in is None:
            self.stage_in = {}
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.stage_ready = torch.cuda.Event()
            self.stage_free = torch.cuda.Event()
            self.stage_free.record(torch.cuda.current_stream(dev))
        self.copy_stream.wait_event(self.stage_free)       # the previous staged batch was consumed
        for k in [k for k in self.stage_in if k not in inputs]:
            del self.stage_in[k]
        with torch.cuda.stream(self.copy_stream):
            for k, v in inputs.items():
                buf = self.stage_in.get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = self.stage_in[k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
                buf.copy_(v, non_blocking=True)
            self.stage_ready.record(self.copy_stream)

    def __call__(self, inputs=None):
        """Run one optimisation step on ``inputs`` (host or device tensors with the example's
        shapes; None = the batch staged by ``prefetch()``); returns the loss as a 0-dim device
        tensor (valid until the next call).  Host inputs are copied asynchronously: do not rewrite
        pinned input tensors before ``self.inputs_consumed`` has completed."""
        if inputs is None:
            assert self.stage_in is not None, 'GraphedTrainStep: nothing was prefetched'
            main = torch.cuda.current_stream(self.trainer.device)
            main.wait_event(self.stage_ready)
            inputs = self.stage_in
            self._load(inputs)
            self.stage_free.record(main)
        else:
            self._load(inputs)
        if self.graph is None:          # capture=False: same data path, eager launches
            self.loss = self._eager()
            return self.loss
        self.graph.replay()
        self.trainer.steps += 1
        return self.loss

    def close(self):
        self.model.kappa_buffers = None
        if self._pinned:
            L.unpin_workspaces()
            self._pinned = False
