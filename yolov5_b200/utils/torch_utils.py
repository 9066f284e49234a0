"""Hot-path edges of the reference's utils/torch_utils.py: BatchNorm folding and the DDP wrapper."""
from __future__ import annotations

import os

import torch
from torch import nn


def fuse_conv_and_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    """Conv2d + eval BatchNorm2d -> one Conv2d with bias; same algebra as reference utils/torch_utils.py:224-254
    (W' = diag(g / sqrt(var + eps)) W,  b' = beta + (b - mean) g / sqrt(var + eps))."""
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                      conv.groups, bias=True).requires_grad_(False).to(conv.weight.device, conv.weight.dtype)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    fused.weight.copy_(conv.weight * scale.view(-1, 1, 1, 1))
    b0 = conv.bias if conv.bias is not None else torch.zeros_like(scale)
    fused.bias.copy_(bn.bias + (b0 - bn.running_mean) * scale)
    return fused


def smart_DDP(model):
    """The path's one collective: gradient all-reduce through DistributedDataParallel over NCCL
    (reference utils/torch_utils.py:61-70)."""
    from torch.nn.parallel import DistributedDataParallel as DDP

    local_rank = int(os.getenv("LOCAL_RANK", -1))
    return DDP(model, device_ids=[local_rank], output_device=local_rank, static_graph=True)


def de_parallel(model):
    return model.module if isinstance(model, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)) else model


class GraphedTrainStep:
    """One optimisation step -- forward (batch-statistics BN), ComputeLoss, backward, gradient clipping, optimizer.step --
    captured once in a CUDA graph and replayed per batch (reference train.py:401-421 is the eager equivalent).

    The training path issues ~500 kernels per step and never synchronises with the host, so the whole step is
    capturable; replaying it removes the Python / launch-issue time that bounds the eager step (yolov5s, 16 images:
    14.6 ms eager -> 10.4 ms replayed on a B200).  Shapes are fixed at construction: `batch` uint8 images of `size` and
    up to `max_targets` label rows; shorter label tensors are padded with zero-size boxes, which build_targets can never
    match (the anchor ratio test of utils/loss.py:219 fails for w = h = 0), so padding does not change the loss.
    Single process only (DDP's bucketed all-reduce hooks are not captured here); optimizers must be capturable
    (torch.optim.SGD is); python-float hyper-parameters (learning rate, momentum) are baked in at capture.

        step = GraphedTrainStep(model, ComputeLoss(model), optimizer, batch=16, size=640)
        for imgs_u8, targets in loader:          # (B,3,H,W) uint8 on any device, (nt,6) float
            loss_items = step(imgs_u8, targets)  # (3,) tensor on the GPU, valid after the replay (stream-ordered)
    """

    def __init__(self, model, compute_loss, optimizer, batch: int, size, max_targets: int | None = None, amp_dtype=torch.float16,
                 max_norm: float | None = 10.0, warmup_steps: int = 3):
        dev = next(model.parameters()).device
        h, w = (size, size) if isinstance(size, int) else size
        self.max_targets = max_targets or 64 * batch
        self.img = torch.zeros(batch, 3, h, w, dtype=torch.uint8, device=dev)
        self.tgt = torch.zeros(self.max_targets, 6, dtype=torch.float32, device=dev)
        self.items = torch.zeros(3, dtype=torch.float32, device=dev)
        params = [p for p in model.parameters() if p.requires_grad]

        def step():
            with torch.autocast("cuda", dtype=amp_dtype):
                pred = model(self.img)
            loss, items = compute_loss(pred, self.tgt)
            optimizer.zero_grad(set_to_none=False)  # gradient tensors keep their addresses across replays
            loss.backward()
            if max_norm is not None:
                torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm)
            optimizer.step()
            self.items.copy_(items)

        # warm-up on a side stream (lazy initialisation, optimizer state, allocator pools), then capture.  The warm-up
        # steps run on the zero batch with a zero learning-rate so they do not move the weights.
        lrs = [g["lr"] for g in optimizer.param_groups]
        state = {k: v.clone() for k, v in model.state_dict().items()}
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for g in optimizer.param_groups:
                g["lr"] = 0.0
            for _ in range(warmup_steps):
                step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for g, lr in zip(optimizer.param_groups, lrs):
            g["lr"] = lr  # python-float hyper-parameters are baked into the capture: rebuild the step to change them
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step()
        # the capture baked in the address of the BN-sum arena (allocated during warm-up, outside the graph's pool): keep it
        # alive even if a later eager forward of another shape makes the arena grow
        from .. import train_ops

        self._arena_buf = train_ops._arena.buf
        # undo what the warm-up touched: BN running statistics / batch counters, momentum buffers
        model.load_state_dict(state)
        for st in optimizer.state.values():
            buf = st.get("momentum_buffer") if isinstance(st, dict) else None
            if buf is not None:
                buf.zero_()

    def __call__(self, imgs_u8: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        nt = targets.shape[0]
        if nt > self.max_targets:
            raise ValueError(f"GraphedTrainStep: {nt} label rows > max_targets {self.max_targets}")
        self.img.copy_(imgs_u8, non_blocking=True)
        self.tgt.zero_()
        if nt:
            self.tgt[:nt].copy_(targets, non_blocking=True)
        self.graph.replay()
        return self.items
