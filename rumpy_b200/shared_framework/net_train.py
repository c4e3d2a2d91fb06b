"""`train_sisr` mirror (reference: rumpy/shared_framework/net_train.py:11-97, training/base_handler.py:206-436):
same command-line options, same TOML schema for the keys the EDSR/RCAN path uses, same on-disk layout
(<save_loc>/<experiment>/{config.toml, saved_models/train_model_<epoch>, result_outputs/summary.csv}).

    python -m rumpy_b200.shared_framework.net_train --parameters cfg.toml [--num_epochs N --gpu single|multi ...]
    torchrun --nproc-per-node 8 -m rumpy_b200.shared_framework.net_train --parameters cfg.toml --gpu multi
"""
import csv
import os

import click


@click.command()
@click.option('--parameters', required=True, help='location of TOML parameters file')
@click.option('--num_epochs', type=int, help='Number of epochs to run through dataset.')
@click.option('--gpu', default=None, type=click.Choice(['single', 'multi'], case_sensitive=False))
@click.option('--sp_gpu', default=None, help='Specify which base GPU to use.')
@click.option('--experiment_name', help='Experiment name to use for saving models/data.')
@click.option('--seed', default=8, show_default=True)
@click.option('--continue_from_epoch', type=int, help='Epoch number from which to resume training.')
@click.option('--overwrite_data', is_flag=True, default=None)
def experiment_setup(parameters, experiment_name, **kwargs):
    import numpy as np
    import toml
    import torch
    import torch.distributed as dist
    from rumpy_b200.shared_framework.data import DevicePairSet, PairSet, psnr_y_device
    from rumpy_b200.shared_framework.models import define_model

    params = toml.load(parameters)
    train_cfg = {**params.get('training', {}), **{k: v for k, v in kwargs.items() if v is not None}}
    if experiment_name is not None:
        params['experiment'] = experiment_name
    seed = int(train_cfg.get('seed', 8))
    torch.manual_seed(seed)
    np.random.seed(seed)
    multi = train_cfg.get('gpu') == 'multi'
    rank, world, local = 0, 1, int(train_cfg.get('sp_gpu') or 0)
    if multi and 'RANK' in os.environ:
        local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        rank, world = dist.get_rank(), dist.get_world_size()
    base = os.path.join(params['experiment_save_loc'], params['experiment'])
    model_dir, out_dir = os.path.join(base, 'saved_models'), os.path.join(base, 'result_outputs')
    if rank == 0:
        os.makedirs(model_dir, exist_ok=True)
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(base, 'config.toml'), 'w') as f:
            toml.dump({**params, 'training': train_cfg}, f)
    internal = dict(params['model'].get('internal_params', {}))
    internal.setdefault('metadata_list', None)      # the reference injects this key (net_train.py:65)
    model = define_model(params['model']['name'], model_save_dir=model_dir, device=local, eval_mode=False,
                         checkpoint_load=None, loss_masking=False, **internal)
    start = 0
    if train_cfg.get('continue_from_epoch') is not None:
        model.load_model('train_model', train_cfg['continue_from_epoch'], legacy=model.legacy_load)
        start = int(train_cfg['continue_from_epoch']) + 1
    if multi:
        model.set_multi_gpu()
    scale = int(internal.get('scale', 4))
    data = params['data']
    # training patches are cut / flipped / converted on the GPU from uint8 images resident in HBM (csrc/glue.cu)
    train_sets = [DevicePairSet(v, scale, seed + rank, device=local) if v.get('crop') else PairSet(v, scale, seed + rank)
                  for v in data['training_sets'].values()]
    eval_sets = [PairSet({**v, 'crop': None, 'random_augment': False}, scale, seed)
                 for v in data.get('eval_sets', {}).values()]
    summary = os.path.join(out_dir, 'summary.csv')
    if rank == 0 and start == 0:
        with open(summary, 'w', newline='') as f:
            csv.writer(f).writerow(['epoch', 'train-loss', 'learning-rate', 'val-loss', 'val-PSNR'])
    for epoch in range(start, int(train_cfg.get('num_epochs', 1))):
        model.set_epoch(epoch)
        losses = []
        for ds in train_sets:
            for batch in ds.batches(int(data['batch_size']), True, 0, 1):
                loss, _ = model.run_train(x=batch['lr'], y=batch['hr'], tag=batch['tag'], keep_on_device=True)
                losses.append(float(loss))
        val_losses, val_psnr = [], []
        if rank == 0:
            for ds in eval_sets:
                for i in range(len(ds)):
                    _, lr, hr = ds.sample(i)
                    hr_dev = hr[None].to(torch.device('cuda', local), non_blocking=True)
                    out, vloss, _ = model.run_eval(lr[None], hr_dev, request_loss=True, keep_on_device=True)
                    val_losses.append(float(vloss))
                    val_psnr.append(float(psnr_y_device(out, hr_dev)[0]))     # clip -> Y -> PSNR on the device
            model.save_model('train_model')
            with open(summary, 'a', newline='') as f:
                csv.writer(f).writerow([epoch, np.mean(losses) if losses else float('nan'), model.get_learning_rate(),
                                        np.mean(val_losses) if val_losses else float('nan'),
                                        np.mean(val_psnr) if val_psnr else float('nan')])
            print(f'epoch {epoch}: train-loss {np.mean(losses):.5f} val-PSNR '
                  f'{np.mean(val_psnr) if val_psnr else float("nan"):.3f}')
        model.epoch_end_calls()
    if multi and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    experiment_setup()
