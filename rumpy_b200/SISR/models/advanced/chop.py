"""Tiled inference (`forward_chop`) shared by the trunk handlers; kept out of handlers.py because the model registry
registers every class it finds there."""
import torch

from rumpy_b200.shared_framework.models.base_architecture import BaseModel


class ChopMixin:
    """Tiled inference with the reference's `forward_chop` semantics (SANHandler.forward_chop,
    /root/reference/rumpy/SISR/models/advanced/handlers.py:85-121): the image is cut into four overlapping quadrants
    (half size + `shave` pixels), each quadrant is super-resolved on its own -- recursively while a quadrant has
    `max_combined_im_size` pixels or more -- and the non-overlapping parts are stitched.  The reference enables this
    for SAN only; here any trunk handler accepts `max_combined_im_size=` (default None = whole image, as the
    reference's EDSR / RCAN / HAN handlers do).  The four quadrants of one level have identical shapes, so they run
    as ONE native batch of 4*b images instead of four calls."""

    max_combined_im_size = None
    scale = 4

    def forward_chop(self, x, shave=10):
        b, c, h, w = x.size()
        h_half, w_half = h // 2, w // 2
        h_size, w_size = h_half + shave, w_half + shave
        lr_list = [
            x[:, :, 0:h_size, 0:w_size],
            x[:, :, 0:h_size, (w - w_size):w],
            x[:, :, (h - h_size):h, 0:w_size],
            x[:, :, (h - h_size):h, (w - w_size):w]]
        if w_size * h_size < self.max_combined_im_size:
            batch = torch.cat([q.contiguous() for q in lr_list], 0)
            sr = BaseModel.run_eval(self, batch, request_loss=False, keep_on_device=True)[0]
            sr_list = list(sr.split(b, 0))
        else:
            sr_list = [self.forward_chop(patch, shave=shave) for patch in lr_list]
        s = self.scale
        h, w = s * h, s * w
        h_half, w_half = s * h_half, s * w_half
        h_size, w_size = s * h_size, s * w_size
        output = sr_list[0].new_empty((b, sr_list[0].shape[1], h, w))
        output[:, :, 0:h_half, 0:w_half] = sr_list[0][:, :, 0:h_half, 0:w_half]
        output[:, :, 0:h_half, w_half:w] = sr_list[1][:, :, 0:h_half, (w_size - w + w_half):w_size]
        output[:, :, h_half:h, 0:w_half] = sr_list[2][:, :, (h_size - h + h_half):h_size, 0:w_half]
        output[:, :, h_half:h, w_half:w] = sr_list[3][:, :, (h_size - h + h_half):h_size, (w_size - w + w_half):w_size]
        return output

    def run_eval(self, x, y=None, request_loss=False, tag=None, timing=False, keep_on_device=False, *args, **kwargs):
        if self.max_combined_im_size is None:
            return BaseModel.run_eval(self, x, y, request_loss=request_loss, tag=tag, timing=timing,
                                      keep_on_device=keep_on_device, *args, **kwargs)
        import time
        tic = time.perf_counter()
        sr = self.forward_chop(x.to(self._torch_device(), non_blocking=True))
        loss = None
        if request_loss and y is not None:
            loss = self.find_loss(sr, y.to(sr.device)).detach().cpu().numpy()
        out = sr if keep_on_device else self._to_host(sr)
        return out, loss, (time.perf_counter() - tic) if timing else None
