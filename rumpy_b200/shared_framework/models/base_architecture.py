"""Mirror of the reference's model-handler base class for the EDSR/RCAN path
(/root/reference/rumpy/shared_framework/models/base_architecture.py:17-612).

Same attribute names, method names, argument meaning, return types and error behaviour as the reference's
`BaseModel`, so `SISRInterface` / `BaseTrainingHandler` / `EvalHub` style callers work unchanged; what differs
is underneath: `self.net` is a rumpy_b200 module whose forward/backward run in librumpy_b200.so, the train step
(L1 loss + backward + Adam) is the fused native step, and `set_multi_gpu` means one-process-per-GPU data
parallelism with an NCCL gradient all-reduce instead of nn.DataParallel (reference :70-77).
"""
import math
import os
import time
from collections import OrderedDict

import numpy as np
import torch
from torch import nn as nn, optim as optim


class BaseModel(nn.Module):
    def __init__(self, device, model_save_dir, eval_mode, grad_clip=None, loss_masking=False, **kwargs):
        super(BaseModel, self).__init__()
        if device == 'cpu' or device == torch.device('cpu'):
            raise RuntimeError("rumpy_b200 handlers need a CUDA (sm_100) device: device='cpu' has no fallback path")
        self.device = device
        self.criterion = nn.L1Loss()            # nn.L1Loss semantics (mean); computed by the native L1 kernel
        self.optimizer = None
        self.net = None
        self.face_finder = False
        self.im_input = None
        self.colorspace = None
        self.steps = None
        self.eval_request_loss = True
        self.loss_masking = loss_masking
        self.grad_clip = None if grad_clip == 0 else grad_clip
        self.model_save_dir = model_save_dir
        self.eval_mode = eval_mode
        self.curr_epoch = 0
        self.state = {}
        self.learning_rate_scheduler = None
        self.legacy_load = True
        self.model_name = self.__class__.__name__.split('Handler')[0].lower()
        self._ddp = None

    # ------------------------------------------------------------------ device / parallelism
    def _torch_device(self):
        d = self.device
        if isinstance(d, torch.device):
            return d
        if isinstance(d, int) or (isinstance(d, str) and d.isnumeric()):
            return torch.device('cuda:%s' % d)
        raise RuntimeError('Device %s not recognized' % d)

    def activate_device(self):
        self.net.to(self._torch_device())

    def set_multi_gpu(self, device_ids=None):
        """Reference :70-77 wraps in nn.DataParallel; here: data parallel = one process per GPU
        (torch.distributed / NCCL), gradients averaged by the bucketed all-reduce in rumpy_b200.parallel."""
        from rumpy_b200 import parallel
        self._ddp = parallel.GradAllReduce(list(self.net.parameters()))
        parallel.broadcast_parameters(self.net)

    # ------------------------------------------------------------------ optimiser / schedulers (reference :79-198)
    def define_optimizer(self, optim_weights, lr=1e-4, optimizer_params=None, optimizer_type='Adam'):
        from rumpy_b200.optim import FusedAdam
        params = [p for p in optim_weights if p.requires_grad]
        if optimizer_type.lower() == 'adam':
            if optimizer_params is not None:
                return FusedAdam(params, lr=lr, betas=(optimizer_params['beta_1'], optimizer_params['beta_2']))
            return FusedAdam(params, lr=lr)
        elif optimizer_type.lower() == 'rmsprop':
            if optimizer_params is not None:
                return optim.RMSprop(params, lr=lr, alpha=optimizer_params['alpha'])
            return optim.RMSprop(params, lr=lr)
        raise RuntimeError('%s optimizer not implemented' % optimizer_type)

    def define_scheduler(self, base_optimizer, scheduler, scheduler_params):
        """Host-side scalar logic only: torch's lr_scheduler objects drive param_groups[0]['lr']."""
        sched = optim.lr_scheduler
        if scheduler == 'cosine_annealing_warm_restarts':
            return sched.CosineAnnealingWarmRestarts(base_optimizer, T_mult=scheduler_params['t_mult'],
                                                     T_0=scheduler_params['restart_period'],
                                                     eta_min=scheduler_params['lr_min'])
        if scheduler == 'one_cycle_lr':
            return sched.OneCycleLR(base_optimizer, max_lr=scheduler_params['lr_max'],
                                    total_steps=scheduler_params['total_steps'],
                                    anneal_strategy=scheduler_params['anneal_strategy'])
        if scheduler == 'multi_step_lr':
            return sched.MultiStepLR(base_optimizer, milestones=scheduler_params['milestones'],
                                     gamma=scheduler_params['gamma'])
        if scheduler == 'step_lr':
            return sched.StepLR(base_optimizer, step_size=scheduler_params['step_size'],
                                gamma=scheduler_params['gamma'])
        if scheduler == 'custom':
            return sched.LambdaLR(base_optimizer, lr_lambda=scheduler_params['function'])
        if scheduler == 'custom_dasr':
            kind = scheduler_params['train_type']
            table = {'long': (60, 225, 100, 125, True), 'short': (21, 79, 35, 44, True),
                     'no_encoder_long': (0, 225, 100, 125, False)}
            if kind not in table:
                raise RuntimeError('Need to select from long or short scheduler type for DASR.')
            warm, flat, off, period, has_warm = table[kind]

            def fn(epoch):
                if has_warm and epoch < warm:
                    return 1e-3
                if epoch < flat:
                    return 1e-4
                return 1e-4 * math.pow(0.5, (epoch - off) // period)
            return sched.LambdaLR(base_optimizer, lr_lambda=fn)
        if scheduler == 'custom_contrastive':
            return sched.LambdaLR(base_optimizer, lr_lambda=lambda it: 0.1 if it < 260 else 5e-4)
        raise RuntimeError('%s scheduler not implemented' % scheduler)

    def training_setup(self, lr, scheduler, scheduler_params, perceptual, device, optimizer_params=None,
                       vgg_type='vgg', vgg_mode='p_loss'):
        if perceptual is not None and self.eval_mode is False:
            raise NotImplementedError('rumpy_b200: perceptual (VGG) loss is outside the EDSR/RCAN trunk path')
        if not self.eval_mode:
            self.optimizer = self.define_optimizer(self.net.parameters(), lr=lr, optimizer_params=optimizer_params)
            if scheduler is not None:
                self.learning_rate_scheduler = self.define_scheduler(self.optimizer, scheduler, scheduler_params)

    # ------------------------------------------------------------------ checkpoints (reference :200-412)
    @staticmethod
    def extract_model_parameters(model):
        return model.state_dict()

    def save_model(self, model_save_name, extract_state_only=False, minimal=False):
        self.state['network'] = self.extract_model_parameters(self.net)
        self.state['model_name'] = self.model_name
        self.state['model_epoch'] = self.curr_epoch
        if not minimal:
            self.state['optimizer'] = self.optimizer.state_dict()
            if self.learning_rate_scheduler is not None:
                self.state['scheduler_G'] = self.learning_rate_scheduler.state_dict()
            if hasattr(self, 'steps'):
                self.state['steps'] = self.steps
        if extract_state_only:
            return self.state
        torch.save(self.state, f=os.path.join(self.model_save_dir, '{}_{}'.format(model_save_name, self.curr_epoch)))

    def load_setup(self, load_override, model_save_name, model_idx):
        loc = str(self._torch_device())
        base = self.model_save_dir if load_override is None else load_override
        return os.path.join(base, '{}_{}'.format(model_save_name, str(model_idx))), loc

    def load_model(self, model_save_name, model_idx, legacy=False, load_override=None, preloaded_state=None,
                   config_changes=None, skip_scheduler_load=False, skip_optimizer_load=False):
        load_file, loc = self.load_setup(load_override, model_save_name, model_idx)
        state = torch.load(f=load_file, map_location=loc, weights_only=False) if preloaded_state is None \
            else preloaded_state
        lr_key = "root['internal_params']['lr']"
        if config_changes is not None and 'values_changed' in config_changes and \
                lr_key in config_changes['values_changed']:
            new_lr = config_changes['values_changed'][lr_key]['new_value']
            for key in [k for k in state.keys() if 'scheduler' in k.lower()]:
                state[key]['base_lrs'] = [new_lr]
                state[key]['_last_lr'] = [new_lr]
            for key in [k for k in state.keys() if 'optimizer' in k.lower()]:
                state[key]['param_groups'][0]['lr'] = new_lr
        net_state = self.legacy_switch(state['network']) if legacy else state['network']
        self.net.load_state_dict(state_dict=net_state)
        if not self.eval_mode:
            if not skip_optimizer_load and 'optimizer' in state:
                self.optimizer.load_state_dict(state['optimizer'])
            if not skip_scheduler_load and self.learning_rate_scheduler is not None and 'scheduler_G' in state:
                self.learning_rate_scheduler.load_state_dict(state['scheduler_G'])
            if hasattr(self, 'steps'):
                self.steps = state.get('steps')
        self.set_epoch(state['model_epoch'])
        print('Loaded model uses the following architecture:', state['model_name'])
        return state

    @staticmethod
    def legacy_switch(state_dict, qrealesrgan_fix=False):
        new_state_dict = OrderedDict()
        for k, v in state_dict.items():
            if k.startswith('model.module.'):
                new_state_dict[k[13:]] = v
            elif k.startswith('model.'):
                new_state_dict[k[6:]] = v
            else:
                new_state_dict[k] = v
        return new_state_dict

    def _copy_stream(self, dev):
        st = getattr(self, '_copy_stream_obj', None)
        if st is None or st.device != dev:
            st = torch.cuda.Stream(device=dev)
            object.__setattr__(self, '_copy_stream_obj', st)
        return st

    @staticmethod
    def _to_host(t):
        """Device -> host through page-locked memory (PyTorch's caching host allocator): a pageable `.cpu()` of the
        SR batch is staged by the driver at a fraction of PCIe bandwidth and dominated the end-to-end step."""
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return out

    # ------------------------------------------------------------------ train / eval (reference :425-520)
    def standard_update(self, loss, scheduler_skip=False):
        """Kept for API parity: with the native trunk the whole update happens inside run_train."""
        raise RuntimeError('rumpy_b200: standard_update is fused into run_train (native L1 + backward + Adam)')

    def run_model(self, x, *args, **kwargs):
        return self.net.forward(x)

    def find_loss(self, out, y):
        from rumpy_b200 import train_native
        return train_native.l1_loss(out, y)

    def get_binary_masks(self, masks):
        raise NotImplementedError('rumpy_b200: loss masking is outside the EDSR/RCAN trunk path')

    def run_train(self, x, y, tag=None, mask=None, keep_on_device=False, scheduler_skip=False, *args, **kwargs):
        if self.eval_mode:
            raise RuntimeError('Model initialized in eval mode, training not possible.')
        if self.loss_masking:
            raise NotImplementedError('rumpy_b200: loss masking is outside the EDSR/RCAN trunk path')
        from rumpy_b200 import train_native
        if not self.net.training:
            self.net.train()
        dev = self._torch_device()
        main = torch.cuda.current_stream(dev)
        copy = self._copy_stream(dev)
        # Host <-> device copies ride a second stream (the reference does H2D -> step -> D2H serially, :473-485):
        #   * x (LR batch, 0.8 MB at 16 x 64 x 64) goes first on the compute stream, the 16x larger HR batch follows on
        #     the copy stream WHILE the forward runs -- it is only needed at the loss;
        #   * the SR batch and the loss are final after the forward, so their device-to-host copies run on the copy
        #     stream underneath the backward + optimiser step.
        x = x.to(device=dev, non_blocking=True)
        if y.device != dev:
            copy.wait_stream(main)                       # orders reuse of the previous step's buffers
            with torch.cuda.stream(copy):
                y = y.to(device=dev, non_blocking=True)
                y_ready = torch.cuda.Event()
                y_ready.record(copy)
        else:
            y_ready = None
        host = {}

        def start_result_copies(loss, out):
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            with torch.cuda.stream(copy):
                copy.wait_event(fwd_done)
                host['loss'] = torch.empty((), dtype=torch.float32, pin_memory=True)
                host['loss'].copy_(loss, non_blocking=True)
                if not keep_on_device:
                    host['out'] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                    host['out'].copy_(out, non_blocking=True)
        loss, out = train_native.train_step(self.net, self.optimizer, x, y, grad_clip=self.grad_clip,
                                            allreduce=self._ddp, metadata=kwargs.get('extra_channels'),
                                            y_ready=y_ready, after_loss=start_result_copies)
        if self.learning_rate_scheduler is not None and not scheduler_skip:
            self.learning_rate_scheduler.step()
        copy.synchronize()                               # results are on the host (the step itself may still run)
        if keep_on_device:
            return host['loss'].numpy().copy(), out.detach()
        main.synchronize()                               # like the reference, the call returns with the step done
        return host['loss'].numpy().copy(), host['out']

    def run_eval(self, x, y=None, request_loss=False, tag=None, timing=False, keep_on_device=False, *args, **kwargs):
        if self.net.training:
            self.net.eval()          # nn.Module.eval() walks ~2000 submodules: only when the mode changes
        with torch.no_grad():
            x = x.to(device=self._torch_device(), non_blocking=True)
            if timing:
                # the reference times perf_counter around forward with no device sync (:504-508);
                # here the forward is bracketed by CUDA events so the number is the device time
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            out = self.run_model(x, image_names=tag, **kwargs)
            if timing:
                e1.record()
            if request_loss and y is not None:
                y = y.to(device=self._torch_device())
                loss = self.find_loss(out, y).detach().cpu().numpy()
            else:
                loss = None
            secs = None
            if timing:
                e1.synchronize()
                secs = e0.elapsed_time(e1) * 1e-3
        if keep_on_device:
            return out.detach(), loss, secs
        return self._to_host(out.detach()), loss, secs

    def run_forensic(self, x, *args, **kwargs):
        raise NotImplementedError('rumpy_b200: forensic() diagnostics are not part of the native trunk')

    # ------------------------------------------------------------------ misc (reference :532-612)
    def print_parameters(self, verbose=False):
        total = 0
        for name, value in self.named_parameters():
            if verbose:
                print(name, value.shape)
            total += int(np.prod(value.shape))
        return total

    def print_status(self):
        raise NotImplementedError

    def epoch_end_calls(self):
        pass

    def extra_diagnostics(self):
        pass

    def pre_training_model_load(self):
        pass

    def verify_eval(self):
        return True

    def set_epoch(self, epoch):
        self.curr_epoch = epoch

    def get_learning_rate(self):
        return self.optimizer.param_groups[0]['lr']

    # direction of improvement per logged metric (reference shared_framework/configuration/constants.py:26-34)
    METRIC_BEST_VAL = {'val-loss': 'lower', 'train-loss': 'lower', 'regression-loss': 'lower',
                       'contrastive-loss': 'lower', 'val-PSNR': 'higher', 'val-SSIM': 'higher', 'val-LPIPS': 'lower'}

    @staticmethod
    def best_model_selection_criteria(log_dir=None, log_file='summary.csv', model_metadata=None, stats=None,
                                      stats_dir=None, base_metric='val-PSNR'):
        """Epoch (row index of summary.csv) with the best `base_metric` (reference base_architecture.py:601-612 ->
        sr_tools/helper_functions.py:29-40: `idxmax` / `idxmin`, i.e. the FIRST best row).  `stats` is a pandas
        DataFrame or a dict of lists (the two forms `load_statistics` returns, sr_tools/stats.py:117-122); when only a
        directory is given the CSV is read from it."""
        if stats is None:
            folder = stats_dir if stats_dir else log_dir
            if folder is None:
                raise RuntimeError('best_model_selection_criteria: neither stats nor a log directory given')
            import csv
            with open(os.path.join(folder, log_file), newline='') as f:
                rows = list(csv.DictReader(f))
            stats = {k: [float(r[k]) for r in rows] for k in rows[0]} if rows else {}
        if base_metric not in BaseModel.METRIC_BEST_VAL:
            raise KeyError(base_metric)
        column = [float(v) for v in list(stats[base_metric])]
        if not column:
            raise RuntimeError(f'best_model_selection_criteria: no rows for {base_metric}')
        arr = np.asarray(column, dtype=np.float64)
        return int(np.nanargmax(arr) if BaseModel.METRIC_BEST_VAL[base_metric] == 'higher' else np.nanargmin(arr))
