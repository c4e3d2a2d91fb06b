"""`eval_sisr` mirror (reference: rumpy/shared_framework/net_eval.py:19-132, evaluation/standard_eval.py:342-556) for
the EDSR/RCAN path: every option of the reference's command is accepted; those that drive other subsystems (face
recognition galleries, dataset splits, metadata files, video) are refused when set.  Images are sharded round-robin over ranks under torchrun (no collective);
rank 0 writes <out_loc>/<results_name>/standard_metrics/{individual,average}_metrics.csv.

    python -m rumpy_b200.shared_framework.net_eval --config eval.toml
"""
import csv
import os

import click

# reference options that are parsed but only valid at their default (value = the reference's default when not falsy)
OUTSIDE_THE_PATH = {'data_attributes': None, 'gallery_source': None, 'galleries': None, 'use_celeba_blacklist': False,
                    'qpi_selection': (None, None), 'gallery_ref_images': None, 'dataset_name': None, 'group_select': None,
                    'image_shortlist': None, 'data_split': None, 'metadata_file': None,
                    'ignore_degradation_location': False, 'augmentation_normalization': None, 'id_source': None,
                    'face_rec_profiling': False, 'save_raw_features': False, 'save_data_model_folders': False,
                    'num_frames': 3, 'hr_selection': 1, 'run_lpips_on_gpu': False}


@click.command()
@click.option('--config', default=None, help='TOML file providing any of the options below')
@click.option('--model_and_epoch', '-me', multiple=True, nargs=2, help='experiment name and epoch')
@click.option('--model_loc', default=None, help='directory holding the experiment folders')
@click.option('--hr_dir', default=None)
@click.option('--lr_dir', default=None)
@click.option('--results_name', default='eval')
@click.option('--out_loc', default='.')
@click.option('--metrics', '-m', multiple=True, default=('PSNR',))
@click.option('--scale', default=4)
@click.option('--batch_size', default=1)
@click.option('--gpu/--no-gpu', default=True)
@click.option('--sp_gpu', default=0)
@click.option('--save_im', is_flag=True, default=False)
@click.option('--full_directory', is_flag=True, help='Ignore data partitions / splits (what this mirror always does).')
@click.option('--no_image_comparison', is_flag=True, help='Accepted: the mirror writes no comparison collages.')
@click.option('--time_models/--no-time_models', default=True, help='Record per-image device time (CUDA events).')
@click.option('--in_features', default=3)
@click.option('--data_type', default='single-frame')
@click.option('--num_image_save', default=100000, help='Stop saving SR images after this many.')
@click.option('--model_only', is_flag=True, help='Skip all metrics and only produce (and save) the SR images.')
@click.option('--lanczos_upsample', is_flag=True)
@click.option('--recursive', default=False)
@click.option('--use_mps', is_flag=True)
# -- options of the reference's evaluation hub that belong to subsystems outside the EDSR / RCAN path (face
#    recognition galleries, dataset splits, metadata-driven models, video): accepted so that existing command lines and
#    TOML files parse, refused with a clear message when they are actually set (net_eval.py:23-58, 71-95)
@click.option('--data_attributes', default=None)
@click.option('--gallery_source', default=None)
@click.option('--galleries', multiple=True, default=None)
@click.option('--use_celeba_blacklist', is_flag=True)
@click.option('--qpi_selection', type=(int, int), default=(None, None))
@click.option('--gallery_ref_images', default=None)
@click.option('--dataset_name', default=None)
@click.option('--group_select', multiple=True, default=None)
@click.option('--image_shortlist', default=None)
@click.option('--data_split', default=None)
@click.option('--metadata_file', default=None)
@click.option('--ignore_degradation_location', is_flag=True)
@click.option('--augmentation_normalization', multiple=True, default=None)
@click.option('--id_source', default=None)
@click.option('--face_rec_profiling', is_flag=True)
@click.option('--save_raw_features', is_flag=True)
@click.option('--save_data_model_folders', is_flag=True)
@click.option('--num_frames', default=3)
@click.option('--hr_selection', default=1)
@click.option('--run_lpips_on_gpu', is_flag=True)
def eval_run(config, **kw):
    import numpy as np
    import toml
    import torch
    import torch.distributed as dist
    from rumpy_b200 import parallel
    from rumpy_b200.shared_framework.data import (PairSet, bicubic_upsample_device, lanczos_upsample_device, psnr_y_device,
                                                  quantize_u8_device)
    from rumpy_b200.shared_framework.models import define_model

    if config:
        for k, v in toml.load(config).items():
            if k in kw:
                kw[k] = v
    if not kw['gpu']:
        raise RuntimeError('rumpy_b200 eval needs a CUDA (sm_100) device: gpu=false has no fallback path')
    outside = [k for k in OUTSIDE_THE_PATH if kw.get(k) not in (None, False, (), [], (None, None), OUTSIDE_THE_PATH[k])]
    if outside:
        raise click.UsageError('rumpy_b200 eval_sisr: option(s) %s belong to subsystems outside the EDSR / RCAN path '
                               '(SURVEY.md section 2); run them through the reference hub with the rumpy_b200 handler '
                               '(INTEGRATION.md section 1)' % ', '.join('--' + k for k in outside))
    if kw['data_type'] != 'single-frame' or int(kw['in_features']) != 3 or kw['use_mps'] or kw['recursive']:
        raise click.UsageError('rumpy_b200 eval_sisr: single-frame RGB images on a CUDA device only')
    metrics = [] if kw['model_only'] else list(kw['metrics'] or [])
    if any(m != 'PSNR' for m in metrics):
        raise click.UsageError('rumpy_b200 eval_sisr: PSNR is the metric computed on the device (SSIM / LPIPS need '
                               'packages absent from this image); got %s' % metrics)
    local = int(kw['sp_gpu'])
    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', 1)) > 1:
        local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    out_dir = os.path.join(kw['out_loc'], kw['results_name'], 'standard_metrics')
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    ds = PairSet({'lr': kw['lr_dir'], 'hr': kw['hr_dir']}, int(kw['scale']))
    rows, saved = [], 0
    # the 'LR' row of the reference's metrics: bicubic baseline of every image (standard_eval.py:371-402), computed on
    # the device with Pillow's own arithmetic (rumpy_bicubic_upsample) and scored next to it
    dev = torch.device('cuda', local)
    for idx in parallel.shard_round_robin(range(len(ds))):
        name, lr, hr = ds.sample(idx)
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        upsample = lanczos_upsample_device if kw['lanczos_upsample'] else bicubic_upsample_device   # standard_eval.py:252
        interp = upsample(lr[None].to(dev), int(kw['scale']))
        t0[1].record()
        score = float(psnr_y_device(interp, hr[None].to(dev))[0]) if metrics else float('nan')
        rows.append({'image': name, 'model': 'LR', 'runtime': t0[0].elapsed_time(t0[1]) * 1e-3 if kw['time_models'] else '',
                     'PSNR': score})
        saved += 1
        if kw['save_im'] and saved <= int(kw['num_image_save']):
            from PIL import Image
            os.makedirs(os.path.join(out_dir, 'bicubic'), exist_ok=True)
            Image.fromarray(quantize_u8_device(interp)[0].cpu().numpy()).save(os.path.join(out_dir, 'bicubic', name))
    for exp, epoch in [tuple(me) for me in kw['model_and_epoch']]:
        cfg = toml.load(os.path.join(kw['model_loc'], exp, 'config.toml'))
        internal = dict(cfg['model'].get('internal_params', {}))
        internal.setdefault('scale', int(kw['scale']))
        internal.pop('lr', None)
        model = define_model(cfg['model']['name'], model_save_dir=os.path.join(kw['model_loc'], exp, 'saved_models'),
                             device=local, eval_mode=True, **internal)
        if epoch in ('best', 'last'):       # reference base_interface.py:86-95: resolved from the training summary
            logs = os.path.join(kw['model_loc'], exp, 'result_outputs')
            if epoch == 'best':
                epoch = model.best_model_selection_criteria(log_dir=logs, base_metric='val-PSNR')
            else:
                with open(os.path.join(logs, 'summary.csv')) as f:
                    epoch = sum(1 for line in f if line.strip()) - 2
        model.load_model('train_model', epoch, legacy=model.legacy_load)
        saved = 0
        for idx in parallel.shard_round_robin(range(len(ds))):
            name, lr, hr = ds.sample(idx)
            # eval glue on the device (csrc/glue.cu): the SR image never crosses PCIe as fp32 -- PSNR(Y) is reduced
            # next to it and only the uint8 image (when saved) and one float come back
            out, _, secs = model.run_eval(lr[None], timing=bool(kw['time_models']), keep_on_device=True)
            hr_dev = hr[None].to(out.device, non_blocking=True)
            rows.append({'image': name, 'model': exp, 'runtime': secs if secs is not None else '',
                         'PSNR': float(psnr_y_device(out, hr_dev)[0]) if metrics else float('nan')})
            saved += 1
            if (kw['save_im'] or kw['model_only']) and saved <= int(kw['num_image_save']):
                from PIL import Image
                im = quantize_u8_device(out)[0].cpu().numpy()      # clip(x*255).astype(uint8): truncation,
                Image.fromarray(im).save(os.path.join(out_dir, f'{exp}_{name}'))                # visualization.py:56
    if dist.is_initialized():
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, rows)      # host-side result collection only; no data-path collective
        rows = [r for part in gathered for r in part]
    if rank == 0:
        with open(os.path.join(out_dir, 'individual_metrics.csv'), 'w', newline='') as f:
            w = csv.DictWriter(f, fieldnames=['image', 'model', 'runtime', 'PSNR'])
            w.writeheader()
            w.writerows(sorted(rows, key=lambda r: (r['model'], r['image'])))
        with open(os.path.join(out_dir, 'average_metrics.csv'), 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['model', 'runtime', 'PSNR'])
            for exp in sorted({r['model'] for r in rows}):
                sel = [r for r in rows if r['model'] == exp]
                times = [r['runtime'] for r in sel if r['runtime'] != '']
                w.writerow([exp, np.mean(times) if times else '', np.mean([r['PSNR'] for r in sel])])
        print(f'wrote {out_dir}')
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    eval_run()
