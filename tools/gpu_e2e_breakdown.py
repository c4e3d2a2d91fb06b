"""Where the host-side time of RCANHandler.run_eval(x_cpu) -> out_cpu goes (16x3x48x48 -> 16x3x192x192)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import recipe
from rumpy_b200.shared_framework.models import define_model

h = define_model('rcan', device=0, model_save_dir=tempfile.mkdtemp(), eval_mode=True, scale=4)
h.net.load_state_dict({k: torch.from_numpy(v) for k, v in recipe.make_weights(recipe.rcan_spec(), seed=8).items()})
x = torch.from_numpy(recipe.make_input((16, 3, 48, 48), 8))
xp = x.pin_memory()
dev = torch.device('cuda:0')


def T(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    print('run_eval(pageable x)      %.3f ms' % T(lambda: h.run_eval(x)))
    print('run_eval(pinned x)        %.3f ms' % T(lambda: h.run_eval(xp)))
    print('run_eval(keep_on_device)  %.3f ms' % T(lambda: h.run_eval(xp, keep_on_device=True)))
    xd = xp.to(dev)
    eng = h.net.native_engine()
    print('forward_graphed (device)  %.3f ms' % T(lambda: eng.forward_graphed(xd)))
    print('forward_inference+clone   %.3f ms' % T(lambda: eng.forward_inference(xd)))
    out = eng.forward_graphed(xd)
    print('_to_host (7 MB, pinned)   %.3f ms' % T(lambda: h._to_host(out)))
    print('h2d pageable              %.3f ms' % T(lambda: x.to(dev, non_blocking=True)))
    print('h2d pinned                %.3f ms' % T(lambda: xp.to(dev, non_blocking=True)))
    buf = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    print('d2h into a reused buffer  %.3f ms' % T(lambda: (buf.copy_(out, non_blocking=True), torch.cuda.current_stream().synchronize())))
    print('net.training check        %.3f ms' % T(lambda: h.net.training))

# effect of bench.py's NVML clock sampler thread on the host-side loop
sys.path.insert(0, ROOT)
import bench
for hz_sleep in (0.1, 0.02):
    s = bench.ClockSampler(0)
    s.run_sleep = hz_sleep
    s.start()
    with torch.no_grad():
        print('run_eval(pinned x) with the NVML sampler thread (bench.py, sleep 0.1 s)   %.3f ms' % T(lambda: h.run_eval(xp)))
    s.result()
    break
