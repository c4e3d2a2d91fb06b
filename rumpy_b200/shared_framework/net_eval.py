"""`eval_sisr` mirror (reference: rumpy/shared_framework/net_eval.py:19-132, evaluation/standard_eval.py:342-556) for
the options the EDSR/RCAN path uses.  Images are sharded round-robin over ranks under torchrun (no collective);
rank 0 writes <out_loc>/<results_name>/standard_metrics/{individual,average}_metrics.csv.

    python -m rumpy_b200.shared_framework.net_eval --config eval.toml
"""
import csv
import os

import click


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
def eval_run(config, **kw):
    import numpy as np
    import toml
    import torch
    import torch.distributed as dist
    from rumpy_b200 import parallel
    from rumpy_b200.shared_framework.data import PairSet, bicubic_upsample_device, psnr_y_device, quantize_u8_device
    from rumpy_b200.shared_framework.models import define_model

    if config:
        for k, v in toml.load(config).items():
            if k in kw:
                kw[k] = v
    if not kw['gpu']:
        raise RuntimeError('rumpy_b200 eval needs a CUDA (sm_100) device: gpu=false has no fallback path')
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
    rows = []
    # the 'LR' row of the reference's metrics: bicubic baseline of every image (standard_eval.py:371-402), computed on
    # the device with Pillow's own arithmetic (rumpy_bicubic_upsample) and scored next to it
    dev = torch.device('cuda', local)
    for idx in parallel.shard_round_robin(range(len(ds))):
        name, lr, hr = ds.sample(idx)
        t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0[0].record()
        interp = bicubic_upsample_device(lr[None].to(dev), int(kw['scale']))
        t0[1].record()
        score = float(psnr_y_device(interp, hr[None].to(dev))[0])
        rows.append({'image': name, 'model': 'LR', 'runtime': t0[0].elapsed_time(t0[1]) * 1e-3, 'PSNR': score})
        if kw['save_im']:
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
        for idx in parallel.shard_round_robin(range(len(ds))):
            name, lr, hr = ds.sample(idx)
            # eval glue on the device (csrc/glue.cu): the SR image never crosses PCIe as fp32 -- PSNR(Y) is reduced
            # next to it and only the uint8 image (when saved) and one float come back
            out, _, secs = model.run_eval(lr[None], timing=True, keep_on_device=True)
            hr_dev = hr[None].to(out.device, non_blocking=True)
            rows.append({'image': name, 'model': exp, 'runtime': secs, 'PSNR': float(psnr_y_device(out, hr_dev)[0])})
            if kw['save_im']:
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
                w.writerow([exp, np.mean([r['runtime'] for r in sel]), np.mean([r['PSNR'] for r in sel])])
        print(f'wrote {out_dir}')
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    eval_run()
