"""Mirror of the reference's Q-RCAN / Q-EDSR handlers
(/root/reference/rumpy/SISR/models/attention_manipulators/handlers.py:11-103): registry names 'qrcan' / 'qedsr',
same constructor arguments and attributes; `scale_qpi` reproduces the 'modulate' style's Gaussian channel scalers."""
import numpy as np
import torch

from rumpy_b200.SISR.models.attention_manipulators import QModel
from rumpy_b200.SISR.models.attention_manipulators.architectures import QEDSR, QHAN, QRCAN


class QRCANHandler(QModel):
    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, scheduler=None,
                 scheduler_params=None, style='modulate', perceptual=None, clamp=False, min_mu=-0.2,
                 max_mu=0.8, n_feats=64, srmd_mode=False, **kwargs):
        super(QRCANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                           **kwargs)
        if srmd_mode or kwargs.get('include_sft_layer'):
            raise NotImplementedError('rumpy_b200 QRCANHandler: SRMD / SFT channel-tiled metadata')
        self.srmd_channel_mode = False
        self.net = QRCAN(scale=scale, in_feats=in_features, num_metadata=self.num_metadata,
                         n_feats=n_feats, style=style, **kwargs)
        self.colorspace = 'augmented_rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'qrcan'
        self.min_mu = min_mu
        self.max_mu = max_mu
        self.base_scaler = np.linspace(0, 1, n_feats)
        self.clamp = clamp
        self.style = style

    @staticmethod
    def gaussian(x, mu, sig=0.2):
        """Normal density N(mu, sig) sampled at `x` (float64 numpy in, fp32 tensor out; reference handlers.py:59-63)."""
        z = (np.asarray(x, dtype=np.float64) - np.asarray(mu, dtype=np.float64)) / sig
        return torch.from_numpy(np.exp(-0.5 * z * z) / (sig * np.sqrt(2.0 * np.pi))).to(torch.float32)

    def scale_qpi(self, qpi):
        """'modulate' style: the [N,1,1,1] quality index becomes an [N,n_feats,1,1] Gaussian bump over the channel
        axis, centred on the index mapped into [min_mu, max_mu] (reference handlers.py:65-73).  One broadcast over
        the batch; float64 density, fp32 result, like the reference."""
        centre = (qpi * (self.max_mu - self.min_mu) + self.min_mu).reshape(-1).numpy()     # fp32, one per image
        bumps = self.gaussian(self.base_scaler[None, :], centre[:, None])
        if self.clamp:
            bumps = bumps.clamp(0, 1)
        return bumps[:, :, None, None]


class QEDSRHandler(QModel):
    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, num_blocks=16,
                 num_features=64, res_scale=0.1, scheduler=None, scheduler_params=None, perceptual=None, **kwargs):
        super(QEDSRHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                           **kwargs)
        self.net = QEDSR(scale=scale, in_features=in_features, num_features=num_features, num_blocks=num_blocks,
                         res_scale=res_scale, input_para=self.num_metadata, **kwargs)
        self.colorspace = 'augmented_rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.model_name = 'qedsr'
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)


class QHANHandler(QModel):
    """reference handlers.py:183-199"""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 scheduler=None, scheduler_params=None, **kwargs):
        super(QHANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                          **kwargs)
        self.net = QHAN(scale=scale, num_metadata=self.num_metadata, **kwargs)
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'qhan'
