"""Mirror of the reference's EDSR / RCAN handlers (/root/reference/rumpy/SISR/models/advanced/handlers.py:8-42):
same registry names ('edsr', 'rcan', 'han'), constructor arguments and attributes."""
from rumpy_b200.SISR.models.advanced.architectures import EDSR, HAN, RCAN
from rumpy_b200.SISR.models.advanced.chop import ChopMixin
from rumpy_b200.shared_framework.models.base_architecture import BaseModel


class EDSRHandler(ChopMixin, BaseModel):
    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, hr_data_loc=None,
                 scheduler=None, scheduler_params=None, perceptual=None,
                 num_features=64, num_blocks=16, res_scale=0.1, max_combined_im_size=None, **kwargs):
        super(EDSRHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                          hr_data_loc=hr_data_loc, **kwargs)
        self.net = EDSR(scale=scale, in_features=in_features, net_features=num_features, num_blocks=num_blocks,
                        res_scale=res_scale)
        self.scale, self.max_combined_im_size = scale, max_combined_im_size
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'edsr'


class RCANHandler(ChopMixin, BaseModel):
    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, perceptual=None,
                 scheduler=None, scheduler_params=None, max_combined_im_size=None, **kwargs):
        super(RCANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                          **kwargs)
        self.net = RCAN(scale=scale, in_feats=in_features, **kwargs)
        self.scale, self.max_combined_im_size = scale, max_combined_im_size
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'rcan'


class HANHandler(ChopMixin, BaseModel):
    """reference handlers.py:44-58 (most parameters locked, as there)."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 scheduler=None, scheduler_params=None, max_combined_im_size=None, **kwargs):
        super(HANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                         **kwargs)
        self.net = HAN(scale=scale)
        self.scale, self.max_combined_im_size = scale, max_combined_im_size
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'han'
