"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python benchmarks/sanitize_target.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from qsparse_b200 import ops, parallel  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
# select: sampled route (n >= 2^22), generic route, batched, ties
v = torch.randn((1 << 22) + 5, device=dev, generator=g)
for k in (0, v.numel() // 2, v.numel() - 1):
    ops.kth_value(v, k)
    ops.kth_value(v, k, take_abs=True)
ops.kth_value(torch.relu(v), v.numel() // 3)
ops.kth_value(v[: 100003], 5000)
ops.kth_value(v[1: 70001], 5000)
vs = [torch.randn(s, device=dev, generator=g) for s in (200_000, 131_072, 999, 300_000)]
ops.kth_value_batched(vs, [s.numel() // 2 for s in vs])
# row-resident kernels: warp per row, CTA per row, TMA variant
for shape in ((33, 8), (64, 288), (130, 1024), (48, 4096), (3, 8200), (2, 16384)):
    x = torch.randn(shape, device=dev, generator=g)
    for kind, w in ((ops.ROW_LINE, 2), (ops.ROW_SCALER, 1), (ops.ROW_DECIMAL, 1)):
        for variant in (0, 1):
            ops.set_tuning(9, variant)
            ops.row_quant_fused_(x, torch.zeros(shape[0], w, device=dev), kind, 4, 1 if kind == ops.ROW_LINE else 0)
ops.set_tuning(9, 0)
# multi-tensor EMA / mask, weight-set step
shapes = [(64, 32, 3, 3), (40, 1000), (7, 13, 5), (256, 64, 3, 3)]
ws = [torch.randn(s, device=dev, generator=g) for s in shapes]
mags = [torch.zeros_like(w) for w in ws]
masks = [torch.ones(w.shape, dtype=torch.bool, device=dev) for w in ws]
outs = [torch.empty_like(w) for w in ws]
for t in range(2):
    parallel.prune_weight_set_step(ws, mags, masks, outs, t, 0.5)
# K9 fused prune step: sampled route (ties, NaN), generic route for the small tensors
xs9 = [torch.randn(s, device=dev, generator=g) * 0.02 for s in (300_000, 131_072 + 5, 999)]
xs9[0] = torch.relu(xs9[0])
xs9[1][::1009] = float("nan")
mg9 = [torch.zeros_like(x) for x in xs9]
mk9 = [torch.ones(x.shape, dtype=torch.bool, device=dev) for x in xs9]
ot9 = [torch.empty_like(x) for x in xs9]
for t in range(2):
    ops.prune_unstructured_step_batched_(mg9, xs9, mk9, ot9, [x.numel() // 2 for x in xs9], t)
# fused training step pieces + maps + export
x = torch.relu(torch.randn(8, 16, 14, 14, device=dev, generator=g))
layout = (8, 16, 196)
mag = torch.zeros(16, device=dev)
mask = torch.ones(16, dtype=torch.bool, device=dev)
scale, dec = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
for t in range(2):
    wsp = ops.reduce_partials(x, layout)
    ops.prune_quant_step_params(mag, mask, scale, dec, wsp, layout, 8 * 196.0, t, 1, t > 0, 8, 8, t, True)
    y = ops.fq_pow2_fwd(x, dec, layout, mask=mask)
    ops.ste_bwd(torch.randn_like(x), dec, True, 8, 0, layout, mask=mask, clamp_in_place=False, want_gx=True)
ops.reduce_stats(x, layout, abssum=True, absmax=True, minmax=True)
ops.fq_line_fwd(x, torch.tensor([[-0.1, 0.9]] * 16, device=dev), 8, True, layout)
ops.fq_scaler_fwd(x, 0.037, (1, 1, x.numel()))
ops.quant_export_int8(x, ops.EXPORT_DECIMAL, 4.0, 8, (1, 1, x.numel()))
ops.quant_export_int8(x, ops.EXPORT_LINE, torch.tensor([[-0.1, 0.9]] * 16, device=dev), 8, layout)
# short-row / channel-last layouts: tile reduction (several visits, ragged tiles, 4-byte-aligned view), transposed
# finalize, row-lane column reduction, group-resident channel-last maps, straddling vectors, packed int4 export
for lay, off in (((300, 24, 196), 0), ((40, 96, 64), 3), ((9, 700, 100), 1), ((2100, 8, 255), 5), ((5000, 24, 1), 0),
                 ((600, 1000, 1), 0), ((64, 96, 49), 0)):
    n_ = lay[0] * lay[1] * lay[2]
    xb = torch.randn(n_ + off, device=dev, generator=g)
    xl = xb[off:].view(lay)
    ops.reduce_stats(xl, lay, abssum=True, absmax=True, minmax=True, nnz=True)
    ops.reduce_stats(xl, lay, minmax=True)
    decl = torch.full((lay[1],), 4.0, device=dev)
    lin = torch.tensor([[-0.5, 0.5]] * lay[1], device=dev)
    ops.fq_pow2_fwd(xl, decl, lay)
    ops.fq_line_fwd(xl, lin, 8, True, lay)
    ops.ste_bwd(xl.clone(), decl, True, 8, 0, lay)
    ops.quant_export_int8(xl, ops.EXPORT_LINE, lin, 4, lay, pack4=True)
    ops.quant_export_int8(xl, ops.EXPORT_SCALER, 0.3, 4, (1, 1, n_), pack4=True)
torch.cuda.synchronize()
print("sanitize target done")
