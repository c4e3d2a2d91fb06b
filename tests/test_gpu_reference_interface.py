"""rumpy_b200 driven by the UNMODIFIED reference's own interface (-m gpu; skipped when baseline/_ref is absent).

INTEGRATION.md section 1 carried out for real: a scratch copy of the installed reference tree (baseline/_ref, written
by tools/install_reference.py) plus the ONE added file `rumpy/SISR/models/b200/handlers.py`.  The reference's registry
scan, `SISRInterface` (SISR/models/interface.py:97-124, shared_framework/models/base_interface.py:30-135), its
checkpoint format and `EvalHub.full_image_protocol` (shared_framework/evaluation/standard_eval.py:342) then run the
sm_100a path under the model names 'rcanb200' / 'edsrb200', next to the reference's own 'rcan' / 'edsr' handlers with
identical weights.  Tolerances are BASELINE.json's: outputs <= 1e-2, loss within 1 %, PSNR within 0.02 dB."""
import csv
import glob
import os
import shutil

import numpy as np
import pytest
import torch

from oracle import ref_import

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(ref_import.INSTALLED, 'rumpy')),
                                 reason='baseline/_ref not installed (tools/install_reference.py)')]

SMALL = {'scale': 4, 'lr': 1e-4, 'n_resgroups': 2, 'n_resblocks': 3}


@pytest.fixture(scope='module')
def ref(tmp_path_factory):
    root = ref_import.overlay_with_b200_handlers(str(tmp_path_factory.mktemp('rumpy_overlay')))
    ref_import.import_reference(root)
    import rumpy.shared_framework.models as registry
    from rumpy.SISR.models.interface import SISRInterface
    assert registry.available_models['rcanb200'] == 'rumpy.SISR.models.b200.handlers.RCANB200Handler'
    assert registry.available_models['rcan'] == 'rumpy.SISR.models.advanced.handlers.RCANHandler'
    return SISRInterface


def _pair(SISRInterface, models, arch, params):
    """(reference handler on the CPU, rumpy_b200 handler on cuda:0) behind two SISRInterface objects, same weights"""
    a = SISRInterface(models, f'{arch}_ref', gpu='off', mode='train', scale=4,
                      new_params={'name': arch, 'internal_params': dict(params)})
    b = SISRInterface(models, f'{arch}_b200', gpu='single', sp_gpu=0, mode='train', scale=4,
                      new_params={'name': arch + 'b200', 'internal_params': dict(params)})
    assert type(a.model).__module__.startswith('rumpy.SISR.models.advanced')
    assert type(b.model).__module__ == 'rumpy.SISR.models.b200.handlers'
    assert list(a.model.net.state_dict().keys()) == list(b.model.net.state_dict().keys())
    b.model.net.load_state_dict(a.model.net.state_dict(), strict=True)
    return a, b


def test_train_eval_and_checkpoints_through_the_reference_interface(ref, tmp_path):
    import toml
    models = str(tmp_path / 'models')
    a, b = _pair(ref, models, 'rcan', SMALL)
    g = torch.Generator().manual_seed(8)
    for step in range(3):
        lr, hr = torch.rand((4, 3, 24, 24), generator=g), torch.rand((4, 3, 96, 96), generator=g)
        la, oa = a.train_batch(lr, hr)
        lb, ob = b.train_batch(lr, hr)
        assert isinstance(lb, np.ndarray) and lb.dtype == np.float32 and ob.device.type == 'cpu'
        assert abs(float(lb) - float(la)) <= 0.01 * float(la), (step, la, lb)
        assert float((oa - ob).abs().max()) <= 1e-2, step
    lr, hr = torch.rand((2, 3, 20, 28), generator=g), torch.rand((2, 3, 80, 112), generator=g)
    rgb_a, y_a, loss_a, _ = a.net_run_and_process(lr=lr, hr=hr, request_loss=True)
    rgb_b, y_b, loss_b, secs = b.net_run_and_process(lr=lr, hr=hr, request_loss=True, timing=True)
    assert rgb_b.shape == rgb_a.shape == (2, 3, 80, 112) and rgb_b.dtype == rgb_a.dtype
    assert float(np.abs(rgb_a - rgb_b).max()) <= 1e-2 and float(np.abs(y_a - y_b).max()) <= 1e-2
    assert abs(float(loss_a) - float(loss_b)) <= 0.01 * float(loss_a) and secs > 0
    # checkpoints travel both ways: same file name, same keys, optimiser state included
    a.model.save_model('train_model')
    b.model.save_model('train_model')
    for exp, name in (('rcan_ref', 'rcan'), ('rcan_b200', 'rcanb200')):
        toml.dump({'model': {'name': name, 'internal_params': SMALL}}, open(os.path.join(models, exp, 'config.toml'), 'w'))
    sa = torch.load(os.path.join(models, 'rcan_ref', 'saved_models', 'train_model_0'), weights_only=False)
    sb = torch.load(os.path.join(models, 'rcan_b200', 'saved_models', 'train_model_0'), weights_only=False)
    assert list(sa.keys()) == list(sb.keys()) and sb['model_name'] == 'rcan'
    assert list(sa['network'].keys()) == list(sb['network'].keys())
    assert sorted(sa['optimizer']['state'].keys()) == sorted(sb['optimizer']['state'].keys())
    for k in ('step', 'exp_avg', 'exp_avg_sq'):
        assert k in sb['optimizer']['state'][0]
    # swap the files: the reference continues from the b200 checkpoint and vice versa
    os.replace(os.path.join(models, 'rcan_ref', 'saved_models', 'train_model_0'), str(tmp_path / 'swap'))
    os.replace(os.path.join(models, 'rcan_b200', 'saved_models', 'train_model_0'),
               os.path.join(models, 'rcan_ref', 'saved_models', 'train_model_0'))
    os.replace(str(tmp_path / 'swap'), os.path.join(models, 'rcan_b200', 'saved_models', 'train_model_0'))
    a2 = ref(models, 'rcan_ref', gpu='off', mode='train', scale=4, load_epoch=0)
    b2 = ref(models, 'rcan_b200', gpu='single', sp_gpu=0, mode='train', scale=4, load_epoch=0)
    la, _ = a2.train_batch(lr, hr)
    lb, _ = b2.train_batch(lr, hr)
    assert abs(float(lb) - float(la)) <= 0.01 * float(la)
    ra, _, _, _ = a2.net_run_and_process(lr=lr)
    rb, _, _, _ = b2.net_run_and_process(lr=lr)
    assert float(np.abs(ra - rb).max()) <= 1e-2


def test_eval_hub_full_image_protocol_on_set5(ref, tmp_path):
    """BASELINE configs[0] through the reference's evaluation hub: EDSR-baseline x4 on the Set5 example images; the
    hub loads both experiments from disk, runs them image by image and writes its metric CSVs."""
    import toml
    from rumpy.shared_framework.evaluation.standard_eval import EvalHub
    models = str(tmp_path / 'models')
    a, b = _pair(ref, models, 'edsr', {'scale': 4, 'lr': 1e-4})
    for itf, exp, name in ((a, 'edsr_ref', 'edsr'), (b, 'edsr_b200', 'edsrb200')):
        itf.model.save_model('train_model')
        toml.dump({'model': {'name': name, 'internal_params': {'scale': 4, 'lr': 1e-4}}},
                  open(os.path.join(models, exp, 'config.toml'), 'w'))
    data = os.path.join(ref_import.INSTALLED, 'Data', 'example_data', 'Set5')
    lr_dir = str(tmp_path / 'lr')            # the PNGs without degradation_metadata.csv (pandas-3 break)
    os.makedirs(lr_dir)
    for f in glob.glob(os.path.join(data, 'lr_random_blur', '*.png')):
        shutil.copy(f, lr_dir)
    hub = EvalHub(**ref_import.eval_hub_kwargs(
        model_and_epoch=[['edsr_ref', '0'], ['edsr_b200', '0']], model_loc=models, hr_dir=os.path.join(data, 'hr'),
        lr_dir=lr_dir, results_name='set5', out_loc=str(tmp_path / 'out'), metrics=['PSNR'], scale=4, batch_size=1,
        full_directory=True, gpu=True, sp_gpu=0, no_image_comparison=True))
    hub.full_image_protocol()
    rows = list(csv.reader(open(str(tmp_path / 'out' / 'set5' / 'standard_metrics' / 'individual_metrics.csv'))))
    models_row, metric_row = rows[0], rows[1]
    col = {(m, k): i for i, (m, k) in enumerate(zip(models_row, metric_row))}
    images = [r for r in rows[2:] if r and r[0].endswith('.png')]
    assert len(images) == 5
    for r in images:
        pa, pb = float(r[col[('edsr_ref', 'PSNR')]]), float(r[col[('edsr_b200', 'PSNR')]])
        assert abs(pa - pb) <= 0.02, (r[0], pa, pb)
        assert float(r[col[('edsr_b200', 'runtime')]]) > 0
