"""torch.profiler breakdown of one ShapeRenderer.forward + backward step (module-level path; see DESIGN.md round-1 numbers)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from tensoflow_b200 import synthetic
from tensoflow_b200.shape_renderer import ShapeRenderer
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
torch.manual_seed(6033)
G0 = 128
cfg = dict(device=dev, gridSize=[G0] * 3, sdf_n_comp=36, sdf_dim=256, app_dim=128, max_levels=1, train_ray_num=4096,
           has_radiance_field=True, radiance_field_step=20000, occ_loss_step=20000, n_samples=64, n_importance=64, up_sample_steps=4)
m = ShapeRenderer(cfg)
for l in (1, 2):
    m.sdf_network.upsample_volume_grid(torch.tensor([G0 << l] * 3))
    m.update_stepSize(torch.tensor([G0 << l] * 3), l + 1)
synthetic.perturb_field(m.sdf_network, seed=1, noise=1e-2)
rays = synthetic.make_rays(4096 * 4, seed=1)
m.set_train_batch({k: v.pin_memory() for k, v in dict(rays_o=rays["rays_o"], rays_d=rays["dirs"], dirs=rays["dirs"], radiis=rays["radiis"],
                                                       rays_cos=rays["rays_cos"], rgbs=rays["rgbs"]).items()})
params = [p for p in m.parameters() if p.requires_grad]
def one_step():
    for p in params: p.grad = None
    out = m({'step': 30000})
    loss = out['loss_rgb'].mean() + 0.1 * out['gradient_error'].mean() + 0.1 * out['loss_tv_sdf']
    for k in ('loss_sparse', 'loss_hessian', 'loss_occ', 'loss_radiance'):
        if k in out and out[k] is not None: loss = loss + 0.01 * out[k].mean()
    loss.backward()
for _ in range(3): one_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    one_step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=45, max_shapes_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))
import time
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): one_step()
t_host = (time.perf_counter() - t) / 5; torch.cuda.synchronize(); t_all = (time.perf_counter() - t) / 5
print(f"host time per step {t_host * 1e3:.1f} ms, wall per step {t_all * 1e3:.1f} ms")
