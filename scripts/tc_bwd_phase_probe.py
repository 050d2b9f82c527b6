"""Timing experiment: stencil backward call with kernel phases disabled (TF_TC_BWD_DEBUG bitmask)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from tensoflow_b200 import ops, synthetic
dev = torch.device('cuda:0')
cfg = dict(bench.SHAPE_CFG); cfg['rays'] = 2048
field, var = bench.build_shape(cfg, dev)
rays = {k: v.to(dev) for k, v in synthetic.make_rays(cfg['rays'], seed=0).items()}
t0, t1, idx = synthetic.uniform_samples(rays['rays_o'], rays['dirs'], field.aabb, cfg['samples'])
mid = (t0 + t1) * 0.5
pts = rays['rays_o'][idx] + rays['dirs'][idx] * mid[:, None]
lv = torch.rand(pts.shape[0], device=dev) * 3 - 0.5
out = field.stencil(pts, lv)
loss = sum((o * torch.randn_like(o)).sum() for o in out if o is not None and o.requires_grad)
for it in range(3):
    ops.KernelTimers.reset(True)
    loss.backward(retain_graph=True)
    torch.cuda.synchronize()
print('TF_TC_BWD_DEBUG', os.environ.get('TF_TC_BWD_DEBUG', '0'), 'samples', pts.shape[0],
      {k: round(v[0], 2) for k, v in ops.KernelTimers.totals_ms().items()})
